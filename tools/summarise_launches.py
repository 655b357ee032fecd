"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total and share."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
kn, mv, mu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    unit = r[mu]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit.startswith("us") else v * 1e3 if unit.startswith("ms") else v)
    name = re.sub(r"\(.*", "", r[kn]).replace("srfrd::", "")
    if "spin_kernel" in name:          # torch.cuda._sleep: bench.py parks the GPU behind it while the host enqueues
        continue
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':40s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:40]:40s} {n:8d} {t:10.1f} {t/n:9.2f} {100*t/tot:5.1f}%")
print(f"{'TOTAL':40s} {sum(a[0] for a in agg.values()):8d} {tot:10.1f}")

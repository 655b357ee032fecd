"""Turn the reports written by tools/profile_round.sh (gpurun_out/<tag>_*) into the committed evidence under profiles/:
<tag>_bench_launches.txt, <tag>_<kernel>_ncu.txt and traffic.json (DRAM bytes per launch, read by bench.py)."""
import csv, json, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "gpurun_out")
prof = os.path.join(root, sys.argv[2]) if len(sys.argv) > 2 else os.path.join(root, "profiles")
os.makedirs(prof, exist_ok=True)


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return [dict(zip(rows[0], r)) for r in rows[2:]], dict(zip(rows[0], rows[1]))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


traffic = {}
try:                                     # keep entries of captures that are not being refreshed
    traffic = json.load(open(os.path.join(root, "profiles", "traffic.json")))
except Exception:
    pass
for name, key, match in [("gemm_tn", "srfrd_gemm_tn", "gemm_tn"), ("topk", "srfrd_catalogue_topk", "catalogue_"),
                         ("k1", "srfrd_embed_ln_fwd@C3", "embed_ln"), ("c3k", "c3k", ""), ("misc", None, "")]:
    rep = os.path.join(out, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep)
        continue
    summary = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(prof, f"{tag}_{name}_ncu.txt"), "w").write(summary)
    if key is None:
        continue
    recs, units = raw(rep)
    per = [to_bytes(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) + to_bytes(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
           for r in recs if match in r["Kernel Name"]]
    if name == "c3k":            # three kernels, one launch each: key per kernel
        for r in recs:
            kn = r["Kernel Name"]
            k2 = ("srfrd_score_loss_fused@C3" if "score_kernel" in kn else "srfrd_embed_bwd@C3" if "embed_bwd" in kn
                  else "srfrd_adam_step@C3" if "adam_kernel" in kn else None)
            if k2:
                traffic[k2] = round(to_bytes(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) +
                                    to_bytes(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"]))
                print(k2, traffic[k2])
        continue
    if name == "topk":           # one pass = streaming kernel + refine kernel
        traffic[key] = round(sum(per))
    else:
        traffic[key] = round(sum(per) / max(len(per), 1))
    print(key, traffic[key], "bytes per launch over", len(per), "launches")
launches = os.path.join(out, f"{tag}_launches.csv")
if os.path.exists(launches):
    s = subprocess.run([sys.executable, os.path.join(root, "tools", "summarise_launches.py"), launches], capture_output=True, text=True).stdout
    open(os.path.join(prof, f"{tag}_bench_launches.txt"), "w").write(s)
    print(s[:1500])
traffic["_source"] = f"ncu --set full --clock-control none ({tag}); dram__bytes_read.sum + dram__bytes_write.sum per launch, cold caches"
json.dump(traffic, open(os.path.join(prof, "traffic.json"), "w"), indent=1)

"""Summarise SRFRD_TOPK_TRACE stamps (catalogue_unitmax_kernel, CTA 0): where a work unit's accumulator round trip goes.
Columns per unit: issuer 0 before / 1 after the wait (tempty + full), 2 after its MMAs, 3 after its commit;
epilogue (quarter 0, column half 0) 4 before / 5 after the wait for tfull, 6 scores loaded + accumulator released,
7 maximum known, 8 after the offer."""
import sys
import numpy as np
b = np.loadtxt(sys.argv[1])
a = b[(b[:, 0] > 0) & (b[:, 5] > 0)][20:-20]
cols = [("issuer: wait tempty+full", 1, 0), ("issuer: MMA issues", 2, 1), ("issuer: commit", 3, 2),
        ("commit -> epilogue sees tfull", 5, 3), ("epilogue: idle before tfull", 5, 4),
        ("epilogue: ld + release", 6, 5), ("epilogue: maximum", 7, 6), ("epilogue: offer", 8, 7),
        ("issuer proceeds -> release", 6, 1)]
for name, hi, lo in cols:
    d = a[:, hi] - a[:, lo]
    print(f"{name:40s} median {np.median(d):7.0f}  mean {d.mean():7.0f}")
t = np.sort(a[:, 2])
print(f"unit period (CTA)                        median {np.median(np.diff(t)):7.0f}  mean {np.diff(t).mean():7.0f}")
per, rel = [], []
for q in range(len(b) - 4):
    if b[q, 1] > 0 and b[q + 4, 1] > 0:
        per.append(b[q + 4, 1] - b[q, 1])
        if b[q, 6] > 0:
            rel.append(b[q + 4, 1] - b[q, 6])
print(f"stage period (same stage, use k -> k+1)  median {np.median(per):7.0f}  mean {np.mean(per):7.0f}")
print(f"release -> same stage's issuer proceeds  median {np.median(rel):7.0f}  mean {np.mean(rel):7.0f}")

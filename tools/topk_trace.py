"""Summarise SRFRD_TOPK_TRACE stamps (catalogue_unitmax_kernel, CTA 0): where a work unit's accumulator round trip goes."""
import sys
import numpy as np
b = np.loadtxt(sys.argv[1])
a = b[(b[:, 0] > 0) & (b[:, 5] > 0)][20:-20]
cols = [("issuer: wait tempty+full", 1, 0), ("issuer: 4 MMA issues", 2, 1), ("issuer: commit", 3, 2),
        ("commit -> epilogue sees tfull", 5, 3), ("epilogue: 64 columns loaded + release", 7, 5),
        ("epilogue: maximum + offer", 8, 7),
        ("epilogue: idle before tfull", 5, 4)]
for name, hi, lo in cols:
    d = a[:, hi] - a[:, lo]
    print(f"{name:40s} median {np.median(d):7.0f}  mean {d.mean():7.0f}")
t = np.sort(a[:, 2])
print(f"unit period (CTA)                        median {np.median(np.diff(t)):7.0f}  mean {np.diff(t).mean():7.0f}")
rt = []
for q in range(len(b) - 3):
    if b[q, 7] > 0 and b[q + 3, 1] > 0:
        rt.append((b[q + 3, 1] - b[q, 7], b[q + 3, 1] - b[q, 1], b[q, 7] - b[q, 1]))
rt = np.array(rt)
print(f"release -> same stage's issuer proceeds  median {np.median(rt[:, 0]):7.0f}")
print(f"stage period (issuer proceeds, use k -> k+1) median {np.median(rt[:, 1]):7.0f}")
print(f"issuer proceeds -> release               median {np.median(rt[:, 2]):7.0f}")

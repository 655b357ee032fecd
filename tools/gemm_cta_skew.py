"""Per-CTA start / duration / SM id of one gemm_tn launch (SRFRD_GEMM_DEBUG=5): where does the CTA skew come from?"""
import ctypes, os, sys
os.environ["SRFRD_GEMM_DEBUG"] = "5"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srfrd_b200 import ops, _lib

T, H = 4096 * 50, 80
bf16 = torch.bfloat16
NB = 6
acts = [torch.randn(T, H, device="cuda").to(bf16) for _ in range(NB)]
outs = [torch.empty(T, H, dtype=bf16, device="cuda") for _ in range(NB)]
res = [torch.randn(T, H, device="cuda").to(bf16) for _ in range(NB)]
W = (torch.randn(H, H, device="cuda") * 0.1).to(bf16)
bias = torch.randn(H, device="cuda")
lib = _lib.load()
buf = (ctypes.c_longlong * (32 * 16 + 320))()
mask48 = (1 << 48) - 1
for rep in range(4):
    for i in range(NB):
        ops.gemm_tn(acts[i], W, out_bf16=outs[i], bias=bias, residual=res[i])
    torch.cuda.synchronize()
    lib.srfrd_gemm_debug_read(buf)
    a = np.array(buf[32 * 16:], dtype=np.int64).reshape(160, 2)[:148]
    start = a[:, 0] & mask48
    end = a[:, 1] & mask48
    smid = (a[:, 1] >> 48) & 0xffff
    t0 = start.min()
    dur = (end - start) / 1e3
    print(f"rep {rep}: kernel span {(end.max() - t0) / 1e3:.1f} us; start spread {(start.max() - t0) / 1e3:.1f} us; "
          f"duration min/med/max {dur.min():.1f}/{np.median(dur):.1f}/{dur.max():.1f} us")
def timeline(tag):
    lib.srfrd_gemm_debug_read(buf)
    d = np.array(buf[:32 * 16], dtype=np.int64).reshape(32, 16)
    c0 = d[10, 1]
    names = ["A issue", "mma tempty", "mma full0", "mma fullL", "epi tfull", "epi buf", "epi math", "epi fence", "epi bar", "epi stored"]
    print(f"--- {tag}: CTA 0 clock64 timeline (cycles since CTA start), per local tile; CTA end {d[11, 1] - c0}")
    print("tile " + " ".join(f"{n:>10s}" for n in names))
    for tl in range(11):
        print(f"{tl:4d} " + " ".join(f"{(d[e, tl] - c0) if d[e, tl] else -1:10d}" for e in range(10)))

for i in range(NB):
    ops.gemm_tn(acts[i], W, out_bf16=outs[i], bias=bias)
torch.cuda.synchronize()
timeline("plain (+bias)")
for i in range(NB):
    ops.gemm_tn(acts[i], W, out_bf16=outs[i], bias=bias, residual=res[i])
torch.cuda.synchronize()
timeline("residual")
order = np.argsort(smid)
print("smid: start_us dur_us  (sorted by smid)")
for k in order:
    print(f"{smid[k]:4d} cta {k:4d} start {(start[k] - t0) / 1e3:6.2f} dur {dur[k]:6.2f} end {(end[k] - t0) / 1e3:6.2f}")

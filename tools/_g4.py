import torch, sys, os, ctypes
os.environ["SRFRD_GEMM_DEBUG"] = "5"
sys.path.insert(0, '.')
from srfrd_b200 import ops, _lib
bf16 = torch.bfloat16
N, K, M = 80, 80, 204800
A = torch.randn(M, K, device='cuda').to(bf16)
B = (torch.randn(N, K, device='cuda') * 0.2).to(bf16)
res = torch.randn(M, N, device='cuda').to(bf16)
bias = torch.randn(N, device='cuda')
o = torch.empty(M, N, dtype=bf16, device='cuda')
use_res = len(sys.argv) > 1
for _ in range(3):
    ops.gemm_tn(A, B, out_bf16=o, bias=bias, residual=res if use_res else None)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
_lib.call("srfrd_gemm_debug_read", ctypes.addressof(buf))
names = ["P:empty0", "M:tempty", "M:full0", "M:fullL", "E:tfull", "E:bar1", "E:math", "E:fence", "E:bar2", "E:end", "x", "x", "M:pre", "M:issued", "M:commit"]
t0 = min(buf[e * 16 + t] for e in (0,1,2,3,4,5,6,7,8,9,12,13,14) for t in range(11) if buf[e * 16 + t] > 0)
print("tile " + " ".join(f"{names[e]:>9s}" for e in (1,2,12,13,14,3,4)))
for t in range(12):
    print(f"{t:4d} " + " ".join(f"{(buf[e * 16 + t] - t0) if buf[e*16+t] > 0 else -1:9d}" for e in (1,2,12,13,14,3,4)))

ns = buf[11 * 16] - buf[10 * 16]; cyc = buf[11 * 16 + 1] - buf[10 * 16 + 1]
print(f"CTA0: {ns} ns, {cyc} cycles -> {cyc / max(ns, 1):.3f} GHz")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.gemm_tn(A, B, out_bf16=o, bias=bias, residual=res if use_res else None)
e1.record(); torch.cuda.synchronize()
print("eager us/launch", e0.elapsed_time(e1) / 20 * 1e3)

"""Per-kernel timing at C2 shapes (T = 4096 x 50 tokens, H = 80): CUDA events over back-to-back launches on
rotating buffers (working set > L2).  Usage: python tools/kbench.py [gemm|attn|ln|wgrad|k1|score|all] [iters]
Under ncu use iters=1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srfrd_b200 import ops

bf16 = torch.bfloat16
what = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B, L, H, D = int(os.environ.get("KB_B", 4096)), int(os.environ.get("KB_L", 50)), int(os.environ.get("KB_H", 80)), 64
T = B * L
NBUF = 6
g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*shape, dtype=bf16, scale=1.0):
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


def timeit(name, fn, bytes_per_call, flops=0.0):
    for i in range(min(3, iters)):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if iters > 1:                       # one CUDA graph of `iters` launches: host launch cost stays out of the timing
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(iters):
                fn(i)
        graph.replay()
        torch.cuda.synchronize()
        e0.record()
        graph.replay()
        e1.record()
    else:
        e0.record()
        fn(0)
        e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    extra = f"  {flops / us / 1e6:7.1f} TFLOP/s" if flops else ""
    print(f"{name:44s} {us:8.1f} us  {bytes_per_call / us / 1e3:7.1f} GB/s{extra}", flush=True)


acts = [rnd(T, H) for _ in range(NBUF)]
outs = [torch.empty(T, H, dtype=bf16, device="cuda") for _ in range(NBUF)]
res = [rnd(T, H) for _ in range(NBUF)]
W = rnd(H, H, scale=0.1)
W2 = rnd(2 * H, H, scale=0.1)
bias = rnd(H, dtype=torch.float32)
bias2 = rnd(2 * H, dtype=torch.float32)
ids = (torch.rand(T, device="cuda") > 0.5).long()
A = T * H * 2

if what in ("gemm", "all"):
    timeit("gemm_tn N=80 K=80 +bias", lambda i: ops.gemm_tn(acts[i % NBUF], W, out_bf16=outs[i % NBUF], bias=bias), 2 * A, 2.0 * T * H * H)
    timeit("gemm_tn N=80 K=80 +bias+residual", lambda i: ops.gemm_tn(acts[i % NBUF], W, out_bf16=outs[i % NBUF], bias=bias, residual=res[i % NBUF]), 3 * A, 2.0 * T * H * H)
    timeit("gemm_tn N=80 +bias+residual+rowmask", lambda i: ops.gemm_tn(acts[i % NBUF], W, out_bf16=outs[i % NBUF], bias=bias, residual=res[i % NBUF], row_ids=ids), 3 * A + T * 8, 2.0 * T * H * H)
    timeit("gemm_tn N=80 +bias+relu", lambda i: ops.gemm_tn(acts[i % NBUF], W, out_bf16=outs[i % NBUF], bias=bias, relu=True), 2 * A, 2.0 * T * H * H)
    timeit("gemm_tn N=80 gate", lambda i: ops.gemm_tn(acts[i % NBUF], W, out_bf16=outs[i % NBUF], gate=res[i % NBUF]), 3 * A, 2.0 * T * H * H)
    kv = [torch.empty(T, 2 * H, dtype=bf16, device="cuda") for _ in range(NBUF)]
    timeit("gemm_tn N=160 K=80 +bias (kv)", lambda i: ops.gemm_tn(acts[i % NBUF], W2, out_bf16=kv[i % NBUF], bias=bias2), 3 * A, 4.0 * T * H * H)
    WT = rnd(H, 2 * H, scale=0.1)
    timeit("gemm_tn N=80 K=160 (dkv -> dx)", lambda i: ops.gemm_tn(kv[i % NBUF], WT, out_bf16=outs[i % NBUF]), 3 * A, 4.0 * T * H * H)

if what in ("wgrad", "all"):
    dW = torch.zeros(H, H, device="cuda")
    db = torch.zeros(H, device="cuda")
    timeit("gemm_wgrad 80x80 (+bias)", lambda i: ops.gemm_wgrad(acts[i % NBUF], res[i % NBUF], dW, db), 2 * A, 2.0 * T * H * H)

if what in ("attn", "all"):
    kv = [rnd(T, 2 * H, scale=0.5) for _ in range(NBUF)]
    timeit("attention_fwd L=%d H=%d" % (L, H), lambda i: ops.attention_fwd(acts[i % NBUF], kv[i % NBUF][:, :H], kv[i % NBUF][:, H:], outs[i % NBUF], B, L, H, 1), 4 * A, 2.0 * B * L * L * H)
    dkv = [torch.empty(T, 2 * H, dtype=bf16, device="cuda") for _ in range(2)]
    timeit("attention_bwd L=%d H=%d" % (L, H), lambda i: ops.attention_bwd(res[i % NBUF], acts[i % NBUF], kv[i % NBUF][:, :H], kv[i % NBUF][:, H:], outs[i % NBUF], dkv[i % 2][:, :H], dkv[i % 2][:, H:], B, L, H, 1), 7 * A, 5.0 * B * L * L * H)

if what in ("ln", "all"):
    w, b = rnd(H, dtype=torch.float32), rnd(H, dtype=torch.float32)
    st = torch.empty(T, 2, device="cuda")
    timeit("layernorm_fwd", lambda i: ops.layernorm_fwd(acts[i % NBUF], w, b, 1e-8, y_bf16=outs[i % NBUF], stats=st), 2 * A + T * 8)
    dw, dbb = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    timeit("layernorm_bwd", lambda i: ops.layernorm_bwd(res[i % NBUF], acts[i % NBUF], st, w, outs[i % NBUF], dw, dbb), 3 * A + T * 8)
    timeit("layernorm_bwd +add+rowmask", lambda i: ops.layernorm_bwd(res[i % NBUF], acts[i % NBUF], st, w, outs[i % NBUF], dw, dbb, add=acts[(i + 1) % NBUF], row_ids=ids), 4 * A + T * 16)

if what in ("k1", "all"):
    N = int(os.environ.get("KB_N", 12101))
    E, P, Fe = rnd(N + 1, 64, dtype=torch.float32), rnd(L, 64, dtype=torch.float32), rnd(3, H - 64, dtype=torch.float32)
    seq = torch.randint(0, N + 1, (B, L), device="cuda")
    valid = float(os.environ.get("KB_VALID", 0.5))
    seq[:, : int(L * (1 - valid))] = 0
    rsq = torch.randint(0, 3, (B, L), device="cuda")
    w, b = rnd(H, dtype=torch.float32), rnd(H, dtype=torch.float32)
    st = torch.empty(T, 2, device="cuda")
    nvalid = int((seq != 0).sum())
    mode = 1 if H > 64 else 0                       # KB_H=64: no fake-review columns (SASRec / SRFU shaped)
    timeit(f"embed_ln_fwd (K1) N={N} valid={nvalid / T:.2f}", lambda i: ops.embed_ln_fwd(E, P, Fe if mode else None, mode, seq, rsq if mode else None, 1.0, w, b, 1e-8, x0_bf16=outs[i % NBUF], q_bf16=acts[i % NBUF], stats=st), 2 * A + T * 24 + nvalid * 256)

if what in ("c3scale",):
    # K4 / K5 / K7 at catalogue scale (for ncu --set full: dram__bytes of one launch each); same shapes as bench.py
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    print(bench.bench_scale_kernels(torch.device("cuda"), bench.peaks()))

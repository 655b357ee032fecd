"""A/B timing of the catalogue kernels (K8) on one GPU: full pass, streaming phase only, ablations, both refine kernels,
at the full catalogue and at an 8-way shard; then an equality check of the two refine kernels and a clock64 trace.
Usage: python tools/topk_ab.py [U] [trace_path]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from srfrd_b200 import evaluation as EV, ops

U = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
trace_path = sys.argv[2] if len(sys.argv) > 2 else None
D = 64
g = torch.Generator().manual_seed(0)
feats = torch.randn(U, D, generator=g).cuda()


def run(idx, fb, u_pad, chunks, ps, pi, packed=None):
    ops.catalogue_topk(fb, U, u_pad, 1, idx.table, idx.row_lo, idx.id_base, chunks, ps, pi, packed)


def timed(idx, env, iters=10):
    for k in ("SRFRD_TOPK_DEBUG", "SRFRD_TOPK_REFINE", "SRFRD_TOPK_ALIGN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    fb, u_pad = EV._split_feats(feats, idx.Dp, 1)
    chunks = ops.catalogue_topk_plan(U, idx.n_rows, idx.row_lo, idx.Dp, 1)
    ps = torch.empty(U, chunks, 10, dtype=torch.float32, device="cuda")
    pi = torch.empty(U, chunks, 10, dtype=torch.int32, device="cuda")
    for _ in range(2):
        run(idx, fb, u_pad, chunks, ps, pi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run(idx, fb, u_pad, chunks, ps, pi)
    e1.record()
    torch.cuda.synchronize()
    for k in env:
        os.environ.pop(k, None)
    return e0.elapsed_time(e1) / iters, ps[:, 0].clone(), pi[:, 0].clone()


for N in (1_000_000, 125_000):
    table = (torch.randn(N + 1, D, generator=g) * 0.05).cuda()
    idx = EV.CatalogueIndex(table, 0)
    base = None
    quick = os.environ.get("AB_QUICK") == "1"
    variants = [("default", {}), ("phase 1 only", {"SRFRD_TOPK_DEBUG": "2"}), ("no insertions", {"SRFRD_TOPK_DEBUG": "1"}),
                ("plain split", {"SRFRD_TOPK_ALIGN": "0"}), ("default again", {})]
    if not quick:
        variants += [("general refine", {"SRFRD_TOPK_REFINE": "0"}), ("one K step", {"SRFRD_TOPK_DEBUG": "4"}),
                     ("no tcgen05.ld", {"SRFRD_TOPK_DEBUG": "5"})]
    for name, env in variants:
        ms, sc, ids = timed(idx, env)
        note = ""
        if name == "default":
            base = (sc, ids)
        elif name == "general refine":
            same = torch.equal(ids, base[1])
            note = f"  ids equal to default: {same}; max |score diff| {float((sc - base[0]).abs().max()):.2e}"
            if not same:
                bad = (ids != base[1]).any(1).nonzero().flatten()
                note += f"  rows differing: {bad.numel()}"
        print(f"N={N:8d} U={U} {name:22s} {ms:7.3f} ms  {U / ms * 1e3 / 1e6:6.2f} M users/s  {2.0 * U * N * D / ms / 1e9:7.1f} TFLOP/s{note}",
              flush=True)
    if N == 1_000_000 and not quick:
        # dyadic data: every accumulation order is exact -> both refine kernels and torch must agree bit for bit
        gd = torch.Generator().manual_seed(5)
        fd = (torch.randint(-16, 17, (512, D), generator=gd).float() / 8).cuda()
        td = (torch.randint(-16, 17, (N + 1, D), generator=gd).float() / 8).cuda()
        idxd = EV.CatalogueIndex(td, 0)
        s1, i1 = EV.local_topk(fd, idxd, 1)
        full = fd[:64] @ td.T
        full[:, 0] = float("-inf")
        # (score desc, id asc): stable sort of -score
        order = torch.sort(-full, dim=1, stable=True).indices[:, :10]
        print("dyadic 1M top-10 ids equal to torch stable sort on 64 users:", torch.equal(order, i1[:64]), flush=True)
        if trace_path:
            os.environ["SRFRD_TOPK_TRACE"] = trace_path
            timed(idx, {}, iters=1)
            os.environ.pop("SRFRD_TOPK_TRACE", None)

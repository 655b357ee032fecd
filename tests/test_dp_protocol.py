"""world_size-2 gloo tests (CPU) of the data-parallel protocol in srfrd_b200/parallel.py, with the CPU oracle
standing in for the CUDA kernels: (1) batch-sharded gradients normalised by the all-reduced weight sums and
SUM-all-reduced equal the single-process gradient; (2) row-sharded top-10 + all-gather + merge equals the
unsharded top-10 bit for bit."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import srfrd_oracle as O
from tests.conftest import load_golden


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2, port=29611):
    torch.set_num_threads(1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _grad_job(rank, world):
    from srfrd_b200 import parallel as P
    fx = load_golden("SRFR")
    batch = dict(fx["in"])
    g = torch.Generator().manual_seed(5)
    p_fake = torch.rand(batch["pos"].shape, generator=g)
    w_full = O.discriminator_weights(batch["pos"], p_fake, "soft")
    shard = P.shard_batch({**batch, "w": w_full}, rank, world)
    tr = O.OracleTrainer(fx["param"], "SRFR", 1)
    _, zp, zn = O.forward(tr.sd, "SRFR", shard["seq"], shard["rsq"], shard["pos"], shard["prs"], shard["neg"], shard["nrs"], 1)
    w = shard["w"]
    W = P.global_weight_sums(torch.stack([w.sum(), w.sum()]))
    loss = (w * torch.nn.functional.softplus(-zp)).sum() / W[0] + (w * torch.nn.functional.softplus(zn)).sum() / W[1]
    loss.backward()
    tr._zero_pad_rows()
    names = sorted(tr.sd)
    flat = torch.cat([(tr.sd[n].grad if tr.sd[n].grad is not None else torch.zeros_like(tr.sd[n])).flatten() for n in names])
    P.allreduce_gradients(flat)
    total = loss.detach().clone()
    dist.all_reduce(total)
    return flat.numpy(), float(total)


def test_sharded_gradients_equal_single_process():
    fx = load_golden("SRFR")
    batch = dict(fx["in"])
    g = torch.Generator().manual_seed(5)
    p_fake = torch.rand(batch["pos"].shape, generator=g)
    w = O.discriminator_weights(batch["pos"], p_fake, "soft")
    tr = O.OracleTrainer(fx["param"], "SRFR", 1)
    ref_loss, ref = tr.grads(batch, w, w)
    names = sorted(ref)
    ref_flat = torch.cat([ref[n].flatten() for n in names]).numpy()
    outs = _run(_grad_job)
    for flat, total in outs:
        assert abs(total - ref_loss) < 1e-5
        np.testing.assert_allclose(flat, ref_flat, rtol=2e-4, atol=2e-6)
    np.testing.assert_array_equal(outs[0][0], outs[1][0])       # identical gradients on every rank


def _topk_job(rank, world):
    from srfrd_b200 import parallel as P
    g = torch.Generator().manual_seed(9)
    feats = torch.randint(-16, 17, (37, 32), generator=g).float() / 8
    table = torch.randint(-16, 17, (1001, 32), generator=g).float() / 8
    lo, hi = P.shard_rows(table.shape[0], rank, world)
    first = 1 if lo == 0 else 0                                   # id 0 is the pad row, never a candidate
    sc = (feats @ table[lo + first:hi].T).numpy()
    v, i = O.topk_stable(sc, 10, first_id=lo + first)
    vs, is_ = P.allgather_topk(torch.from_numpy(v), torch.from_numpy(i))
    mv, mi = O.merge_topk(vs.reshape(37, -1).numpy(), is_.reshape(37, -1).numpy(), 10)
    return mi


def test_sharded_topk_equals_unsharded():
    g = torch.Generator().manual_seed(9)
    feats = torch.randint(-16, 17, (37, 32), generator=g).float() / 8
    table = torch.randint(-16, 17, (1001, 32), generator=g).float() / 8
    _, ref = O.catalogue_topk(feats, table, 10)
    outs = _run(_topk_job, port=29612)
    for mi in outs:
        np.testing.assert_array_equal(mi, ref)


def test_shard_bounds_partition_the_table():
    from srfrd_b200 import parallel as P
    from srfrd_b200.evaluation import CatalogueIndex
    for n in (1, 7, 12102, 1_000_001):
        for G in (1, 2, 4, 8):
            covered = 0
            for r in range(G):
                lo, hi = P.shard_rows(n, r, G)
                assert (lo, hi) == CatalogueIndex.shard_bounds(n, r, G)
                assert lo == covered or lo == hi == n
                covered = max(covered, hi)
            assert covered == n

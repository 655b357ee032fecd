"""world_size-2 gloo tests (CPU) of the data-parallel protocol in srfrd_b200/parallel.py -- the SAME functions
FusedTrainer._step_body and evaluation.sharded_topk call on the GPU (allreduce_sum_ on the weight sums and on the
gradient bucket whose tail carries the loss accumulators; allgather_packed_topk on the 80-byte wire format) -- with the
CPU oracle standing in for the CUDA kernels: (1) batch-sharded gradients normalised by the all-reduced weight sums and
SUM-all-reduced in one bucket equal the single-process gradient and loss; (2) row-sharded top-10 + ONE packed
all-gather + merge equals the unsharded top-10 bit for bit."""
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
    W = P.allreduce_sum_(torch.stack([w.sum(), w.sum()]), dist.group.WORLD)      # trainer: norm on the comm stream
    acc = torch.stack([(w * torch.nn.functional.softplus(-zp)).sum(), (w * torch.nn.functional.softplus(zn)).sum()])
    loss = acc[0] / W[0] + acc[1] / W[1]
    loss.backward()
    tr._zero_pad_rows()
    names = sorted(tr.sd)
    flat = torch.cat([(tr.sd[n].grad if tr.sd[n].grad is not None else torch.zeros_like(tr.sd[n])).flatten() for n in names])
    bucket = torch.cat([flat, acc.detach(), torch.zeros(2)])                     # FlatParams.grad_bucket: grads + 4-float tail
    P.allreduce_sum_(bucket, dist.group.WORLD)                                   # trainer: ONE collective per step
    total = bucket[-4] / W[0] + bucket[-3] / W[1]                                # loss_finalize
    return bucket[:-4].numpy(), float(total)


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
    # the scoring kernel's wire format: 10 fp32 scores + 10 int32 global ids per user (80 B), one all-gather
    packed = torch.from_numpy(np.concatenate([v.astype(np.float32), i.astype(np.int32).view(np.float32)], 1))
    gathered = P.allgather_packed_topk(packed, dist.group.WORLD)
    assert gathered.shape == (world, 37, 20)
    mv, mi = P.merge_packed_topk_host(gathered, 10)
    mv2, mi2 = O.merge_topk(gathered[:, :, :10].permute(1, 0, 2).reshape(37, -1).numpy(),
                            gathered[:, :, 10:].contiguous().view(torch.int32).permute(1, 0, 2).reshape(37, -1).numpy().astype(np.int64), 10)
    assert np.array_equal(mi, mi2)
    return mi


def test_sharded_topk_equals_unsharded():
    g = torch.Generator().manual_seed(9)
    feats = torch.randint(-16, 17, (37, 32), generator=g).float() / 8
    table = torch.randint(-16, 17, (1001, 32), generator=g).float() / 8
    _, ref = O.catalogue_topk(feats, table, 10)
    outs = _run(_topk_job, port=29612)
    for mi in outs:
        np.testing.assert_array_equal(mi, ref)


class _RowwiseEncoder:
    """Stands in for the model: every user's representation depends on that user's row only (as the encoder's does)."""

    def __init__(self):
        g = torch.Generator().manual_seed(3)
        self.w = torch.randn(50, 80, generator=g)

    def encode_last(self, seq, rsq):
        return torch.tanh(seq.float() @ self.w) + (0 if rsq is None else rsq.float().sum(1, keepdim=True))


def _encode_job(rank, world):
    from srfrd_b200 import evaluation as EV
    g = torch.Generator().manual_seed(11)
    seq = torch.randint(0, 9, (37, 50), generator=g)                # 37 users over 2 ranks: 19 + 18, one padded row
    rsq = torch.randint(0, 3, (37, 50), generator=g)
    return EV.encode_users(_RowwiseEncoder(), seq, rsq, dist.group.WORLD).numpy()


def test_user_sharded_encode_equals_the_replicated_one():
    """evaluation.encode_users: users split over the ranks + ONE all-gather == every rank encoding everybody."""
    g = torch.Generator().manual_seed(11)
    seq = torch.randint(0, 9, (37, 50), generator=g)
    rsq = torch.randint(0, 3, (37, 50), generator=g)
    ref = _RowwiseEncoder().encode_last(seq, rsq).numpy()
    for out in _run(_encode_job, port=29613):
        np.testing.assert_array_equal(out, ref)


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

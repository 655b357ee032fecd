"""Host-side data-parallel protocol (one process per GPU, torch.distributed; NCCL over NVLink on the GPU box).

Training: the global batch is split by rows, parameters and Adam state are replicated, and
    loss = sum_ranks [ sum_{t in rank} w_t * bce_t ] / W        with  W = all_reduce(sum_t w_t)
so every rank divides by the GLOBAL weight sum (the reference's mean over ``pos != 0`` of the whole batch,
trainer.py:36-38) and gradients are SUM-all-reduced: G ranks on B/G sequences reproduce one rank on B.
FusedTrainer implements exactly this on the device (two tiny all-reduces for W and the loss, one for the flat
gradient buffer).

Catalogue scoring: the (N+1)-row item table is row-sharded, every rank scores all users against its rows, and the
per-rank top-10 lists (80 B per user) are all-gathered and merged with the tie-break (score desc, id asc).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of rank `rank` when n rows are split over `world` ranks (ceil split)."""
    per = (n + world - 1) // world
    return min(rank * per, n), min((rank + 1) * per, n)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Rank r takes rows [r*B/G, (r+1)*B/G) of every (B, ...) tensor in the batch."""
    B = next(iter(batch.values())).shape[0]
    lo, hi = shard_rows(B, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def global_weight_sums(local_sums: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce the per-rank (sum w_pos, sum w_neg) pair so each rank normalises by the global sums."""
    out = local_sums.clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def allreduce_gradients(flat_grad: torch.Tensor, group=None) -> None:
    """One SUM all-reduce over the flat fp32 gradient buffer (3.4 MB at C2)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)


def broadcast_parameters(flat_params: torch.Tensor, src: int = 0, group=None) -> None:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat_params, src, group=group)


def allgather_topk(scores: torch.Tensor, ids: torch.Tensor, group=None):
    """(U, 10) per rank -> (U, G, 10) candidate lists on every rank."""
    G = dist.get_world_size(group)
    gs = [torch.empty_like(scores) for _ in range(G)]
    gi = [torch.empty_like(ids) for _ in range(G)]
    dist.all_gather(gs, scores.contiguous(), group=group)
    dist.all_gather(gi, ids.contiguous(), group=group)
    return torch.stack(gs, 1), torch.stack(gi, 1)

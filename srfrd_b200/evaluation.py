"""Evaluation on the hot path: full-catalogue top-k (tcgen05 scoring fused with a streaming top-10)
and the reference's sampled-101 protocol (utils.py:544-602), both batched.

Full-catalogue scoring is the reference's ``predict(user_ids, seq, rsq, label)`` called with
``label = arange(1, itemnum + 1)`` followed by ``(-logits).argsort().argsort()`` (utils.py:589-591),
i.e. a ranking by (score desc, item id asc) -- except that the (U, N) logits are never written to HBM.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import ops

bf16 = torch.bfloat16
TK = 10


class CatalogueIndex:
    """bf16 copy of (a row shard of) the item table, resident on the device for scoring.

    table_f32 : (n_rows_local, D) fp32 rows of the item table owned by this rank
    id_base   : global item id of local row 0 (0 on the rank that owns the pad row)
    Item id 0 is the pad row and is never a candidate (utils.py:575-583 draws candidates from 1..itemnum).
    """

    def __init__(self, table_f32: torch.Tensor, id_base: int = 0):
        self.id_base = int(id_base)
        self.n_rows, self.D = table_f32.shape
        self.Dp = (self.D + 15) // 16 * 16          # zero-padded K (tcgen05 K step / 16-byte rows)
        self.row_lo = 1 if self.id_base == 0 else 0
        self.table = torch.zeros(self.n_rows, self.Dp, dtype=bf16, device=table_f32.device)
        ops.f32_to_bf16_split(table_f32.contiguous(), self.table[:, :self.D], None)

    @staticmethod
    def shard_bounds(n_rows_total: int, rank: int, world: int) -> Tuple[int, int]:
        """Row-sharding of the (N+1)-row table: rank g owns rows [g*S, min((g+1)*S, N+1)), S = ceil((N+1)/G)."""
        S = (n_rows_total + world - 1) // world
        return min(rank * S, n_rows_total), min((rank + 1) * S, n_rows_total)


def local_topk(feats_f32: torch.Tensor, index: CatalogueIndex, n_split: int = 1):
    """Per-user top-10 of feats (U, D) against the local shard -> (scores (U,10) fp32, ids (U,10) int64)."""
    U, D = feats_f32.shape
    dev = feats_f32.device
    Dp = index.Dp
    u_pad = (U + 255) // 256 * 256          # whole user groups (2 tiles of 128) per split
    fb = torch.zeros(n_split * u_pad, Dp, dtype=bf16, device=dev)
    f = feats_f32.contiguous()
    if n_split == 1:
        ops.f32_to_bf16_split(f, fb[:U, :D], None)
    elif n_split == 2:
        ops.f32_to_bf16_split(f, fb[:U, :D], fb[u_pad:u_pad + U, :D])
    else:
        # three-way split: hi, mid, lo  (f = hi + mid + lo to ~24 mantissa bits)
        hi, mid = fb[:U, :D], fb[u_pad:u_pad + U, :D]
        ops.f32_to_bf16_split(f, hi, mid)
        rest = f - hi.float() - mid.float()
        ops.f32_to_bf16_split(rest.contiguous(), fb[2 * u_pad:2 * u_pad + U, :D], None)
    if index.n_rows <= index.row_lo:                      # empty shard
        return (torch.full((U, TK), float("-inf"), device=dev), torch.full((U, TK), -1, dtype=torch.int64, device=dev))
    chunks = ops.catalogue_topk_plan(U, index.n_rows, index.row_lo, Dp, n_split)
    ps = torch.empty(U, chunks, TK, dtype=torch.float32, device=dev)
    pi = torch.empty(U, chunks, TK, dtype=torch.int32, device=dev)
    ops.catalogue_topk(fb, U, u_pad, n_split, index.table, index.row_lo, index.id_base, chunks, ps, pi)
    out_s = torch.empty(U, TK, dtype=torch.float32, device=dev)
    out_i = torch.empty(U, TK, dtype=torch.int64, device=dev)
    ops.merge_topk(ps, pi, U, chunks, TK, out_s, out_i)
    return out_s, out_i


def merge_shards(scores: torch.Tensor, ids: torch.Tensor, k: int = TK):
    """scores/ids: (U, G, 10) candidate lists gathered from G shards -> global (U, k)."""
    U, G, _ = scores.shape
    out_s = torch.empty(U, k, dtype=torch.float32, device=scores.device)
    out_i = torch.empty(U, k, dtype=torch.int64, device=scores.device)
    ops.merge_topk(scores.contiguous(), ids.to(torch.int32).contiguous(), U, G, k, out_s, out_i)
    return out_s, out_i


def sharded_topk(feats_f32: torch.Tensor, index: CatalogueIndex, process_group=None, n_split: int = 1):
    """Row-sharded catalogue scoring: local top-10, all-gather (80 B / user / rank), merge -> identical
    (U, 10) on every rank."""
    s, i = local_topk(feats_f32, index, n_split)
    if process_group is None or torch.distributed.get_world_size(process_group) == 1:
        return s, i
    G = torch.distributed.get_world_size(process_group)
    U = s.shape[0]
    gs = torch.empty(G, U, TK, dtype=torch.float32, device=s.device)
    gi = torch.empty(G, U, TK, dtype=torch.int64, device=s.device)
    torch.distributed.all_gather_into_tensor(gs, s.contiguous(), group=process_group)
    torch.distributed.all_gather_into_tensor(gi, i.contiguous(), group=process_group)
    return merge_shards(gs.permute(1, 0, 2).contiguous(), gi.permute(1, 0, 2).contiguous())


def hr_ndcg_from_topk(topk_ids: torch.Tensor, target: torch.Tensor, k: int = 10) -> Tuple[float, float]:
    """(NDCG@k, HR@k) in the reference's return order (utils.py:595-602): a hit at 0-based rank r < k adds
    1 to HT and 1/log2(r+2) to NDCG; means over users.  Reads k ids per user back to the host."""
    ids = topk_ids[:, :k].cpu().numpy()
    tgt = np.asarray(target.cpu() if torch.is_tensor(target) else target).reshape(-1, 1)
    hit = ids == tgt
    rank = hit.argmax(1)
    has = hit.any(1)
    ndcg = np.where(has, 1.0 / np.log2(rank + 2.0), 0.0)
    n = max(len(tgt), 1)
    return float(ndcg.sum() / n), float(has.sum() / n)


@torch.no_grad()
def evaluate_full_catalogue(model, seq, rsq, target, batch_users: int = 16384, n_split: int = 1, process_group=None,
                            index: Optional[CatalogueIndex] = None):
    """Encode users, score the whole catalogue, return (NDCG@10, HR@10, topk_ids)."""
    eng = model._sync_flat()
    if index is None:
        index = CatalogueIndex(eng.P.view(model.spec.item_key), 0)
    outs = []
    for s in range(0, seq.shape[0], batch_users):
        feats = model.encode_last(seq[s:s + batch_users], None if rsq is None else rsq[s:s + batch_users])
        _, ids = sharded_topk(feats[:, :model.spec.D], index, process_group, n_split)
        outs.append(ids)
    ids = torch.cat(outs)
    ndcg, hr = hr_ndcg_from_topk(ids, target)
    return ndcg, hr, ids


@torch.no_grad()
def evaluation(model, dataset, maxlen, device, max_users: int = 10000, seed: Optional[int] = None, chunk: int = 2048):
    """Batched restatement of the reference's evaluation() (utils.py:544-602): per user 1 held-out target +
    100 uniform negatives not in the user's train set, rank of the target among the 101, HR@10 / NDCG@10.
    Returns (NDCG@10, HR@10).  The candidate scoring goes through model.predict's tensor-core path."""
    train, test, usernum, itemnum = dataset
    rng = np.random.default_rng(seed)
    users = list(range(1, usernum + 1))
    if usernum > max_users:
        users = rng.choice(np.arange(1, usernum + 1), max_users, replace=False).tolist()
    users = [u for u in users if len(train["item_ids"][u]) >= 1 and len(test["item_ids"][u]) >= 1]
    NDCG = HT = 0.0
    for s in range(0, len(users), chunk):
        us = users[s:s + chunk]
        seq = np.zeros((len(us), maxlen), np.int64)
        rsq = np.zeros((len(us), maxlen), np.int64)
        cand = np.zeros((len(us), 101), np.int64)
        for r, u in enumerate(us):
            it, rv = train["item_ids"][u][-maxlen:], train["review_ids"][u][-maxlen:]
            seq[r, maxlen - len(it):] = it
            rsq[r, maxlen - len(rv):] = rv
            rated = set(train["item_ids"][u]) | {0}
            cand[r, 0] = test["item_ids"][u][0]
            j = 1
            while j < 101:
                t = int(rng.integers(1, itemnum + 1))
                if t not in rated:
                    cand[r, j] = t
                    j += 1
        feats = model.encode_last(torch.from_numpy(seq).to(device), torch.from_numpy(rsq).to(device))
        table = model._engine.P.view(model.spec.item_key)
        rows = table[torch.from_numpy(cand).to(device)]                     # (u, 101, D) candidate gather
        logits = torch.einsum("ud,ucd->uc", feats[:, :model.spec.D], rows)   # 101 dots per user: not the hot path
        rank = (logits[:, 1:] > logits[:, :1]).sum(1).cpu().numpy()
        hit = rank < 10
        NDCG += float(np.where(hit, 1.0 / np.log2(rank + 2.0), 0.0).sum())
        HT += float(hit.sum())
    n = max(len(users), 1)
    return NDCG / n, HT / n

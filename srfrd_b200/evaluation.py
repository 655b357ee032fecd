"""Evaluation on the hot path: full-catalogue top-k (tcgen05 scoring fused with a streaming top-10)
and the reference's sampled-101 protocol (utils.py:544-602), both batched and both on the library's own kernels.

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
        if self.n_rows > 0:                           # (more ranks than table rows: an empty shard is legal)
            ops.f32_to_bf16_split(table_f32.contiguous(), self.table[:, :self.D], None)

    @staticmethod
    def shard_bounds(n_rows_total: int, rank: int, world: int) -> Tuple[int, int]:
        """Row-sharding of the (N+1)-row table: rank g owns rows [g*S, min((g+1)*S, N+1)), S = ceil((N+1)/G)."""
        S = (n_rows_total + world - 1) // world
        return min(rank * S, n_rows_total), min((rank + 1) * S, n_rows_total)


def _split_feats(feats_f32: torch.Tensor, Dp: int, n_split: int):
    U, D = feats_f32.shape
    dev = feats_f32.device
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
    return fb, u_pad


def _local_lists(feats_f32: torch.Tensor, index: CatalogueIndex, n_split: int, packed: Optional[torch.Tensor]):
    """Run K8 on the local shard; returns the (U, chunks, 10) list buffers whose slot 0 holds each user's exact local
    top-10 (sorted, global ids) and, when `packed` is given, the same list in the all-gather wire format."""
    U = feats_f32.shape[0]
    dev = feats_f32.device
    fb, u_pad = _split_feats(feats_f32, index.Dp, n_split)
    chunks = ops.catalogue_topk_plan(U, index.n_rows, index.row_lo, index.Dp, n_split)
    ps = torch.empty(U, chunks, TK, dtype=torch.float32, device=dev)
    pi = torch.empty(U, chunks, TK, dtype=torch.int32, device=dev)
    ops.catalogue_topk(fb, U, u_pad, n_split, index.table, index.row_lo, index.id_base, chunks, ps, pi, packed)
    return ps, pi, chunks


def local_topk(feats_f32: torch.Tensor, index: CatalogueIndex, n_split: int = 1):
    """Per-user top-10 of feats (U, D) against the local shard -> (scores (U,10) fp32, ids (U,10) int64)."""
    U = feats_f32.shape[0]
    dev = feats_f32.device
    if index.n_rows <= index.row_lo:                      # empty shard
        return (torch.full((U, TK), float("-inf"), device=dev), torch.full((U, TK), -1, dtype=torch.int64, device=dev))
    ps, pi, chunks = _local_lists(feats_f32, index, n_split, None)
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
    """Row-sharded catalogue scoring: local top-10, ONE all-gather of the packed lists (10 fp32 scores + 10 int32 global
    ids = 80 B / user / rank, written in wire format by the scoring kernel itself), merge straight out of the gathered
    buffer -> identical (U, 10) on every rank.  No int64 on the wire, no permute / contiguous copies."""
    from . import parallel
    if process_group is None or torch.distributed.get_world_size(process_group) == 1:
        return local_topk(feats_f32, index, n_split)
    G = torch.distributed.get_world_size(process_group)
    U = feats_f32.shape[0]
    dev = feats_f32.device
    packed = torch.empty(U, 2 * TK, dtype=torch.float32, device=dev)
    if index.n_rows <= index.row_lo:                      # empty shard: ids -1 (bit pattern of int32 -1 in every id word)
        packed.view(torch.int32).fill_(-1)
    else:
        _local_lists(feats_f32, index, n_split, packed)
    gathered = parallel.allgather_packed_topk(packed, process_group)        # (G, U, 20)
    out_s = torch.empty(U, TK, dtype=torch.float32, device=dev)
    out_i = torch.empty(U, TK, dtype=torch.int64, device=dev)
    ops.merge_topk_packed(gathered, U, G, TK, out_s, out_i)
    return out_s, out_i


@torch.no_grad()
def encode_users(model, seq, rsq, process_group=None) -> torch.Tensor:
    """hidden[:, -1, :] for every user (U, Dout) fp32.  With a process group the USERS are split over the ranks (a sequence's
    encoding does not depend on the rest of the batch) and ONE all-gather of the representations (Dout fp32 per user)
    gives every rank all of them -- the item table is sharded by rows, so every rank needs every user, but nobody needs to
    encode them G times (at 8 ranks the replicated encode was 40 % of a catalogue pass)."""
    from . import parallel
    if process_group is None or torch.distributed.get_world_size(process_group) == 1:
        return model.encode_last(seq, rsq)
    G, r = torch.distributed.get_world_size(process_group), torch.distributed.get_rank(process_group)
    U = seq.shape[0]
    per = (U + G - 1) // G
    lo, hi = min(r * per, U), min((r + 1) * per, U)
    if hi > lo:
        mine = model.encode_last(seq[lo:hi], None if rsq is None else rsq[lo:hi])
        width = mine.shape[1]
    else:                                                 # more ranks than users: an empty slice still joins the all-gather
        width = model.encode_last(seq[:1], None if rsq is None else rsq[:1]).shape[1]
        mine = torch.empty(0, width, dtype=torch.float32, device=seq.device if torch.is_tensor(seq) else None)
    buf = torch.zeros(per, width, dtype=torch.float32, device=mine.device)
    buf[:hi - lo] = mine
    return parallel.allgather_rows(buf, process_group)[:U]


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
        feats = encode_users(model, seq[s:s + batch_users], None if rsq is None else rsq[s:s + batch_users], process_group)
        _, ids = sharded_topk(feats[:, :model.spec.D], index, process_group, n_split)
        outs.append(ids)
    ids = torch.cat(outs)
    ndcg, hr = hr_ndcg_from_topk(ids, target)
    return ndcg, hr, ids


def dataset_to_csr(dataset):
    """The reference's [user_train, user_test, usernum, itemnum] dicts (utils.py:92-139) -> CSR arrays over user rows
    0..usernum-1 (row u-1 = user u): offsets int64, items int32, labels int8, first held-out item (0 = none)."""
    train, test, usernum, itemnum = dataset
    lens = np.fromiter((len(train["item_ids"].get(u, ())) for u in range(1, usernum + 1)), np.int64, usernum)
    offsets = np.zeros(usernum + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    items = np.fromiter((i for u in range(1, usernum + 1) for i in train["item_ids"].get(u, ())), np.int32, int(offsets[-1]))
    labels = np.fromiter((r for u in range(1, usernum + 1) for r in train["review_ids"].get(u, ())), np.int8, int(offsets[-1]))
    target = np.fromiter((test["item_ids"][u][0] if len(test["item_ids"].get(u, ())) >= 1 else 0
                          for u in range(1, usernum + 1)), np.int32, usernum)
    return offsets, items, labels, target, int(usernum), int(itemnum)


def eval_users(csr, max_users: int = 10000, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """User rows the reference would evaluate (utils.py:551-559): at most ``max_users`` sampled without replacement,
    then those with at least one train and one held-out interaction."""
    offsets, _, _, target, usernum, _ = csr
    rows = np.arange(usernum)
    if usernum > max_users:
        rng = rng or np.random.default_rng()
        rows = np.sort(rng.choice(usernum, max_users, replace=False))
    ok = (np.diff(offsets)[rows] >= 1) & (target[rows] != 0)
    return rows[ok].astype(np.int32)


def right_aligned(csr, rows: np.ndarray, maxlen: int):
    """seq / rsq as evaluation() builds them (utils.py:561-574): the last ``maxlen`` train items, right-aligned."""
    offsets, items, labels = csr[0], csr[1], csr[2]
    a, b = offsets[rows], offsets[rows + 1]
    t = np.arange(maxlen)[None, :]
    src = (b - a)[:, None] - (maxlen - t)
    valid = src >= 0
    gi = np.where(valid, a[:, None] + src, 0)
    if len(items) == 0:
        z = np.zeros((len(rows), maxlen), np.int64)
        return z, z.copy()
    return (np.where(valid, items[gi], 0).astype(np.int64), np.where(valid, labels[gi], 0).astype(np.int64))


@torch.no_grad()
def sampled_ranks(model, csr, rows: np.ndarray, maxlen: int, device, seed: int = 0, n_neg: int = 100,
                  candidates: Optional[np.ndarray] = None, chunk: int = 8192, return_logits: bool = False):
    """0-based rank of each user's held-out item among itself + ``n_neg`` sampled negatives (utils.py:576-591), on the
    device: candidate draw (srfrd_sample_candidates, rejection against the user's train row), encoder on the hot-path
    kernels, candidate scoring + rank (srfrd_candidate_rank).  ``candidates`` (len(rows), 1 + n_neg) overrides the
    draw (parity tests feed the reference's own sets).  Returns ranks (numpy int64)[, logits (numpy)]."""
    offsets, items, labels, target, usernum, itemnum = csr
    dev = torch.device(device)
    eng = model._sync_flat()
    spec = model.spec
    d_off = torch.from_numpy(offsets).to(dev)
    d_items = torch.from_numpy(items if len(items) else np.zeros(1, np.int32)).to(dev)
    d_target = torch.from_numpy(target).to(dev)
    C = 1 + n_neg if candidates is None else candidates.shape[1]
    ranks, logits_out = [], []
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    for s in range(0, len(rows), chunk):
        r = rows[s:s + chunk]
        seq, rsq = right_aligned(csr, r, maxlen)
        seq_d, rsq_d = torch.from_numpy(seq).to(dev), torch.from_numpy(rsq).to(dev)
        if candidates is None:
            cand = torch.empty(len(r), C, dtype=torch.int64, device=dev)
            ops.sample_candidates(d_off, d_items, torch.from_numpy(np.ascontiguousarray(r, np.int32)).to(dev), d_target,
                                  itemnum, C, seed + s, cand)
        else:
            cand = torch.from_numpy(np.ascontiguousarray(candidates[s:s + chunk], np.int64)).to(dev)
        feats = model.encode_last(seq_d, rsq_d)                          # (u, Dout) fp32
        rank = torch.empty(len(r), dtype=torch.int32, device=dev)
        lg = torch.empty(len(r), C, dtype=torch.float32, device=dev) if return_logits else None
        ft = lab = None
        if spec.kind == "SRFRN":                                         # rows E[id] || Fe[user label], SRFR_model.py:244-257
            ft = eng.P.view("embedding_layer.fake_embed.weight")
            lab = torch.empty(len(r), dtype=torch.int64, device=dev)
            ops.srfu_labels(rsq_d.contiguous(), 3, lab)
        ops.candidate_rank(feats, eng.P.view(spec.item_key), cand, spec.D, ft, lab, lg, rank, err)
        ranks.append(rank)
        if return_logits:
            logits_out.append(lg)
    if not ranks:
        return (np.zeros(0, np.int64), np.zeros((0, C), np.float32)) if return_logits else np.zeros(0, np.int64)
    out = torch.cat(ranks).cpu().numpy().astype(np.int64)                # one D2H read for all users
    if int(err.item()):
        raise IndexError("index out of range in self: a candidate item id is outside the item table")
    if return_logits:
        return out, torch.cat(logits_out).cpu().numpy()
    return out


@torch.no_grad()
def evaluation(model, dataset, maxlen, device, max_users: int = 10000, seed: Optional[int] = None, chunk: int = 8192,
               candidates: Optional[np.ndarray] = None, users: Optional[np.ndarray] = None):
    """The reference's evaluation() (utils.py:544-602), batched on the device: per user 1 held-out target + 100 uniform
    negatives not in the user's train set, rank of the target among the 101, HR@10 / NDCG@10.  Returns
    (NDCG@10, HR@10).  Candidate draw, encoder and candidate scoring / ranking all run on the library's kernels
    (sampled_ranks); ``users`` (1-based ids) / ``candidates`` pin the user order and candidate sets for parity tests."""
    csr = dataset_to_csr(dataset)
    rng = np.random.default_rng(seed)
    rows = eval_users(csr, max_users, rng) if users is None else (np.asarray(users, np.int64) - 1).astype(np.int32)
    rank = sampled_ranks(model, csr, rows, maxlen, device, seed=int(rng.integers(1 << 31)), candidates=candidates, chunk=chunk)
    hit = rank < 10
    n = max(len(rows), 1)
    return float(np.where(hit, 1.0 / np.log2(rank + 2.0), 0.0).sum() / n), float(hit.sum() / n)

"""Host-side data helpers with the reference's ``utils.py`` names (next rows 8f #2, #3).

``df_data_partition``      vectorised restatement of utils.py:92-139 (the reference walks the frame with
                           ``DataFrame.iterrows``, minutes at 10 M interactions); same return value
                           ``[user_train, user_test, usernum, itemnum]`` with the same dict layout.
``interactions_from_df``   the same leave-one-out split as CSR arrays (srfrd_b200.synth.Interactions) for the
                           on-device sampler (trainer.DeviceSampler) and the batched evaluators.
``evaluation_with_label``  batched restatement of utils.py:628-752: sampled-101 ranking plus the per-label breakdown
                           (binary / frequency / ratio user labels, utils.py:604-626); the encoder runs on
                           the device kernels.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .synth import Interactions


def _split_arrays(df, is_valid: bool):
    """Per-interaction arrays in frame order, grouped by user with the frame order kept inside a user."""
    u = np.asarray(df["user_id"].to_numpy(), np.int64)
    it = np.asarray(df["item_id"].to_numpy(), np.int64)
    lab = np.where(df["fake_review"].to_numpy() == "fake", 1, 2).astype(np.int8)        # utils.py:116-121
    order = np.argsort(u, kind="stable")
    u, it, lab = u[order], it[order], lab[order]
    users, start, cnt = np.unique(u, return_index=True, return_counts=True)
    final = 2 if is_valid else 1                                                         # final_idx = -2 / -1
    return u, it, lab, users, start, cnt, final


def df_data_partition(df, is_valid: bool = False):
    """utils.py:92-139.  Users with fewer than 2 interactions keep everything in train and have an empty test
    entry; otherwise train = interactions[:final_idx] and test = [interactions[final_idx]]."""
    u, it, lab, users, start, cnt, final = _split_arrays(df, is_valid)
    user_train = {"item_ids": {}, "review_ids": {}}
    user_test = {"item_ids": {}, "review_ids": {}}
    items_l, labs_l = it.tolist(), lab.tolist()
    # dict insertion order of the reference = order of first appearance in the frame
    first_seen = np.asarray(df["user_id"].to_numpy(), np.int64)
    _, first_idx = np.unique(first_seen, return_index=True)
    for k in np.argsort(first_idx, kind="stable"):
        user, a, n = int(users[k]), int(start[k]), int(cnt[k])
        if n < 2:
            user_train["item_ids"][user] = items_l[a:a + n]
            user_train["review_ids"][user] = labs_l[a:a + n]
            user_test["item_ids"][user] = []
            user_test["review_ids"][user] = []
        else:
            user_train["item_ids"][user] = items_l[a:a + n - final]
            user_train["review_ids"][user] = labs_l[a:a + n - final]
            user_test["item_ids"][user] = [items_l[a + n - final]]
            user_test["review_ids"][user] = [labs_l[a + n - final]]
    usernum = int(u.max()) if len(u) else 0
    itemnum = int(it.max()) if len(it) else 0
    return [user_train, user_test, usernum, itemnum]


def interactions_from_df(df, is_valid: bool = False, p_fake: Optional[np.ndarray] = None) -> Interactions:
    """Same split as df_data_partition, as CSR arrays over users 1..usernum (users absent from the frame get empty
    rows).  ``p_fake`` (per frame row, optional) carries discriminator probabilities for the weighted loss."""
    u, it, lab, users, start, cnt, final = _split_arrays(df, is_valid)
    usernum = int(u.max()) if len(u) else 0
    itemnum = int(it.max()) if len(it) else 0
    if p_fake is None:
        pf = np.where(lab == 1, 1.0, 0.0).astype(np.float32)
    else:
        pf = np.asarray(p_fake, np.float32)[np.argsort(np.asarray(df["user_id"].to_numpy(), np.int64), kind="stable")]
    tr_len = np.zeros(usernum, np.int64)
    has_test = cnt >= 2
    tr_len[users - 1] = np.where(has_test, cnt - final, cnt)
    offsets = np.zeros(usernum + 1, np.int64)
    np.cumsum(tr_len, out=offsets[1:])
    pos_in_user = np.arange(len(u)) - np.repeat(start, cnt)
    keep = pos_in_user < np.repeat(np.where(has_test, cnt - final, cnt), cnt)
    test_item = np.zeros(usernum, np.int32)
    test_label = np.zeros(usernum, np.int8)
    ti = start + cnt - final
    test_item[users[has_test] - 1] = it[ti[has_test]]
    test_label[users[has_test] - 1] = lab[ti[has_test]]
    return Interactions(offsets, it[keep].astype(np.int32), lab[keep], pf[keep], test_item, test_label, usernum, itemnum)


def label_breakdown(ranks: np.ndarray, labels: np.ndarray) -> Dict[int, list]:
    """utils.py:722-752: per user label -> [HT@10, NDCG@10, number of users], sorted by label."""
    hit = ranks < 10
    ndcg = np.where(hit, 1.0 / np.log2(ranks + 2.0), 0.0)
    out = {}
    for lab in np.unique(labels):
        m = labels == lab
        n = int(m.sum())
        out[int(lab)] = [float(hit[m].sum() / n), float(ndcg[m].sum() / n), n]
    return dict(sorted(out.items()))


@torch.no_grad()
def evaluation_with_label(model, dataset, maxlen, device, max_users: int = 10000, seed: Optional[int] = None,
                          chunk: int = 8192, candidates: Optional[np.ndarray] = None, users: Optional[np.ndarray] = None):
    """utils.py:628-752, batched on the device kernels (evaluation.sampled_ranks).  Returns (NDCG@10, HT@10, userResults,
    Binary_Metric, Frequency_Metric, Ratio_Metric); userResults[user] = [rank, HIT, NDCG, label_B, label_F, label_R]."""
    from . import evaluation as EV
    csr = EV.dataset_to_csr(dataset)
    rng = np.random.default_rng(seed)
    rows = EV.eval_users(csr, max_users, rng) if users is None else (np.asarray(users, np.int64) - 1).astype(np.int32)
    if len(rows) == 0:
        return 0.0, 0.0, {}, {}, {}, {}
    rank = EV.sampled_ranks(model, csr, rows, maxlen, device, seed=int(rng.integers(1 << 31)), candidates=candidates,
                            chunk=chunk)
    # user labels of utils.py:604-626.  NOTE the binary rule here (1 = mostly fake) is the INVERSE of
    # SRFU_B.get_Labels (SRFR_model.py:546-552, 2 = mostly fake): both are reproduced as written.
    _, rsq = EV.right_aligned(csr, rows, maxlen)
    nf, nr = (rsq == 1).sum(1), (rsq == 2).sum(1)
    lb = np.where(nf > nr, 1, 2)
    lf = nf
    lr = np.floor(nf / np.maximum(nf + nr, 1) * 10).astype(np.int64)
    hit = rank < 10
    ndcg = np.where(hit, 1.0 / np.log2(rank + 2.0), 0.0)
    user_results = {int(u) + 1: [int(rank[i]), float(hit[i]), float(ndcg[i]), int(lb[i]), int(lf[i]), int(lr[i])]
                    for i, u in enumerate(rows)}
    n = len(rows)
    return (float(ndcg.sum() / n), float(hit.sum() / n), user_results, label_breakdown(rank, lb),
            label_breakdown(rank, lf), label_breakdown(rank, lr))

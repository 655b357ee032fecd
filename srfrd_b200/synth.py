"""Seeded synthetic interaction data shaped like the reference's datasets.

The reference reads a CSV (user_id, item_id, time, fake_review, ...) through
``df_data_partition`` (utils.py:92-139) into per-user python lists and samples batches with
``sample_function_fr`` (utils.py:21-65).  Here the same information lives in a CSR layout
(offsets, items, labels, p_fake) and batches are produced by a vectorised sampler with the
reference's layout: right-aligned, left-padded (id 0) arrays of length ``maxlen``;
``seq`` = items[:-1], ``pos`` = items shifted by one, ``neg`` = uniform item not in the user's
set wherever pos != 0, ``rsq``/``prs`` = the discriminator labels likewise ({1 fake, 2 real}),
``nrs`` == 1 wherever valid (utils.py:52 ``randint(1, 2)``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np


@dataclass
class Interactions:
    """Leave-one-out split in CSR form.  Users are 1..usernum (row u-1), items 1..itemnum."""
    offsets: np.ndarray      # (usernum+1,) int64 into items/labels/p_fake  (TRAIN part)
    items: np.ndarray        # (nnz,) int32
    labels: np.ndarray       # (nnz,) int8   1 = fake, 2 = real   (utils.py:116-121)
    p_fake: np.ndarray       # (nnz,) float32 discriminator probability (extension, row L)
    test_item: np.ndarray    # (usernum,) int32 held-out last item, 0 = none (utils.py:131-136)
    test_label: np.ndarray   # (usernum,) int8
    usernum: int
    itemnum: int

    def train_len(self) -> np.ndarray:
        return np.diff(self.offsets)

    def to_reference_dataset(self):
        """[user_train, user_test, usernum, itemnum] exactly as df_data_partition returns it."""
        tr = {"item_ids": {}, "review_ids": {}}
        te = {"item_ids": {}, "review_ids": {}}
        for u in range(1, self.usernum + 1):
            a, b = self.offsets[u - 1], self.offsets[u]
            tr["item_ids"][u] = self.items[a:b].tolist()
            tr["review_ids"][u] = self.labels[a:b].tolist()
            te["item_ids"][u] = [int(self.test_item[u - 1])] if self.test_item[u - 1] else []
            te["review_ids"][u] = [int(self.test_label[u - 1])] if self.test_item[u - 1] else []
        return [tr, te, self.usernum, self.itemnum]


def _zipf_items(rng, n, itemnum, a):
    # Zipf(a) over 1..itemnum by inverse-CDF on a truncated power law
    w = 1.0 / np.arange(1, itemnum + 1, dtype=np.float64) ** a
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    perm = rng.permutation(itemnum) + 1          # popularity rank -> item id
    return perm[np.searchsorted(cdf, rng.random(n))].astype(np.int32)


def make_interactions(seed: int, usernum: int, itemnum: int, min_len: int, mean_extra: float,
                      max_len: int, fake_rate: float = 0.15, zipf_a: float = 1.0,
                      lognormal: bool = False) -> Interactions:
    """Per-user total length = min_len + Geometric(mean mean_extra) (or lognormal), clipped to
    max_len+1 so train length <= max_len; items Zipf(zipf_a) without repeats inside a user."""
    rng = np.random.default_rng(seed)
    if lognormal:
        tot = np.clip(np.round(rng.lognormal(np.log(mean_extra) - 0.5, 1.0, usernum)), min_len, max_len + 1)
    else:
        tot = np.clip(min_len + rng.geometric(1.0 / max(mean_extra, 1.0), usernum) - 1, min_len, max_len + 1)
    tot = tot.astype(np.int64)
    offs_all = np.zeros(usernum + 1, np.int64)
    np.cumsum(tot, out=offs_all[1:])
    nnz = int(offs_all[-1])
    items = _zipf_items(rng, nnz, itemnum, zipf_a)
    user_of = np.repeat(np.arange(usernum, dtype=np.int64), tot)
    # remove within-user repeats by re-drawing uniformly (a few rounds suffice)
    for _ in range(8):
        key = user_of * (itemnum + 1) + items
        order = np.argsort(key, kind="stable")
        dup = np.zeros(nnz, bool)
        dup[order[1:]] = key[order[1:]] == key[order[:-1]]
        if not dup.any():
            break
        items[dup] = rng.integers(1, itemnum + 1, int(dup.sum()), dtype=np.int32)
    fake = rng.random(nnz) < fake_rate
    labels = np.where(fake, 1, 2).astype(np.int8)
    p = np.where(fake, rng.normal(0.8, 0.1, nnz), rng.normal(0.2, 0.1, nnz))
    p_fake = np.clip(p, 0.0, 1.0).astype(np.float32)
    # leave-one-out: last interaction of every user with >= 2 is the test item (utils.py:124-136)
    last = offs_all[1:] - 1
    has_test = tot >= 2
    test_item = np.where(has_test, items[last], 0).astype(np.int32)
    test_label = np.where(has_test, labels[last], 0).astype(np.int8)
    keep = np.ones(nnz, bool)
    keep[last[has_test]] = False
    tr_len = tot - has_test
    offsets = np.zeros(usernum + 1, np.int64)
    np.cumsum(tr_len, out=offsets[1:])
    return Interactions(offsets, items[keep], labels[keep], p_fake[keep], test_item, test_label, usernum, itemnum)


CONFIGS: Dict[str, dict] = {
    # SURVEY.md 8d.  D/F/L/blocks are model hyper-parameters; the rest shapes the data.
    "C1": dict(seed=1235, usernum=22363, itemnum=12101, min_len=5, mean_extra=4.0, max_len=50,
               D=64, F=16, L=50, blocks=2, heads=1, batch=128),
    "C2": dict(seed=1236, usernum=22363, itemnum=12101, min_len=5, mean_extra=4.0, max_len=50,
               D=64, F=16, L=50, blocks=2, heads=1, batch=4096),
    "C3": dict(seed=1237, usernum=16384, itemnum=1_000_000, min_len=5, mean_extra=4.0, max_len=50,
               D=64, F=16, L=50, blocks=2, heads=1, batch=16384),
    "C4": dict(seed=1238, usernum=22363, itemnum=12101, min_len=20, mean_extra=60.0, max_len=200,
               D=256, F=16, L=200, blocks=4, heads=1, batch=1024),
    "C5": dict(seed=1239, usernum=500_000, itemnum=150_000, min_len=3, mean_extra=24.5, max_len=50,
               D=64, F=16, L=50, blocks=2, heads=1, batch=4096, fake_rate=0.30, lognormal=True),
}


def make_config_data(name: str, usernum: Optional[int] = None, itemnum: Optional[int] = None) -> Interactions:
    c = CONFIGS[name]
    return make_interactions(c["seed"], usernum or c["usernum"], itemnum or c["itemnum"], c["min_len"],
                             c["mean_extra"], c["max_len"], c.get("fake_rate", 0.15), 1.0,
                             c.get("lognormal", False))


class BatchSampler:
    """Vectorised restatement of sample_function_fr (utils.py:21-65).  Users with <= 1 train item are
    rejected (utils.py:25).  Emits int64 arrays (the reference casts to LongTensor, trainer.py:29)."""

    def __init__(self, data: Interactions, maxlen: int, seed: int = 0):
        self.d, self.L = data, maxlen
        self.rng = np.random.default_rng(seed)
        self.eligible = np.nonzero(data.train_len() > 1)[0]
        lens = data.train_len()
        user_of = np.repeat(np.arange(data.usernum, dtype=np.int64), lens)
        self._keys = np.sort(user_of * (data.itemnum + 1) + data.items)

    def _in_user_set(self, users0: np.ndarray, items: np.ndarray) -> np.ndarray:
        key = users0.astype(np.int64) * (self.d.itemnum + 1) + items
        j = np.searchsorted(self._keys, key)
        j = np.minimum(j, len(self._keys) - 1)
        return self._keys[j] == key

    def next_batch(self, batch_size: int) -> Dict[str, np.ndarray]:
        d, L = self.d, self.L
        u0 = self.rng.choice(self.eligible, batch_size)               # 0-based user rows
        a, b = d.offsets[u0], d.offsets[u0 + 1]
        n = b - a                                                     # train length >= 2
        t = np.arange(L)[None, :]
        # slot t holds train index (n-1) - (L - t) for seq and +1 for pos; valid if >= 0
        src = (n[:, None] - 1) - (L - t)
        valid = src >= 0
        gi = np.where(valid, a[:, None] + src, 0)
        seq = np.where(valid, d.items[gi], 0).astype(np.int64)
        rsq = np.where(valid, d.labels[gi], 0).astype(np.int64)
        pos = np.where(valid, d.items[np.where(valid, gi + 1, 0)], 0).astype(np.int64)
        prs = np.where(valid, d.labels[np.where(valid, gi + 1, 0)], 0).astype(np.int64)
        pfk = np.where(valid, d.p_fake[np.where(valid, gi + 1, 0)], 0.0).astype(np.float32)
        neg = self.rng.integers(1, d.itemnum + 1, (batch_size, L))
        uu = np.broadcast_to(u0[:, None], neg.shape)
        for _ in range(64):
            bad = self._in_user_set(uu, neg) & valid
            if not bad.any():
                break
            neg[bad] = self.rng.integers(1, d.itemnum + 1, int(bad.sum()))
        neg = np.where(valid, neg, 0).astype(np.int64)
        nrs = valid.astype(np.int64)
        return dict(u=(u0 + 1).astype(np.int64), seq=seq, rsq=rsq, pos=pos, prs=prs, neg=neg, nrs=nrs,
                    p_fake=pfk)


def eval_sequences(data: Interactions, maxlen: int, users0: np.ndarray):
    """Right-aligned (seq, rsq) over the TRAIN items of the given users, as evaluation() builds them
    (utils.py:561-574), plus the held-out target item."""
    a, b = data.offsets[users0], data.offsets[users0 + 1]
    n = b - a
    t = np.arange(maxlen)[None, :]
    src = n[:, None] - (maxlen - t)
    valid = src >= 0
    gi = np.where(valid, a[:, None] + src, 0)
    seq = np.where(valid, data.items[gi], 0).astype(np.int64)
    rsq = np.where(valid, data.labels[gi], 0).astype(np.int64)
    return seq, rsq, data.test_item[users0].astype(np.int64)

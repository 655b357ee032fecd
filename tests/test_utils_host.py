"""Host-side data helpers against the reference's own functions (golden produced by oracle/make_partition_golden.py
from the unmodified utils.py:92-139) and against hand restatements of utils.py:604-626 / :722-752."""
import json
import os

import numpy as np
import pandas as pd

from tests.conftest import GOLDEN


def _fx():
    return json.load(open(os.path.join(GOLDEN, "partition.json")))


def test_df_data_partition_matches_reference():
    from srfrd_b200.utils import df_data_partition
    fx = _fx()
    df = pd.DataFrame(fx["frame"])
    for valid in (False, True):
        tr, te, un, inum = df_data_partition(df, valid)
        ex = fx["expect"][str(valid)]
        assert un == ex["usernum"] and inum == ex["itemnum"]
        assert [int(k) for k in tr["item_ids"].keys()] == ex["order"]           # same dict order as the reference
        for k in ex["order"]:
            assert tr["item_ids"][k] == ex["train_items"][str(k)] and tr["review_ids"][k] == ex["train_reviews"][str(k)]
            assert te["item_ids"][k] == ex["test_items"][str(k)] and te["review_ids"][k] == ex["test_reviews"][str(k)]


def test_interactions_from_df_is_the_same_split_in_csr_form():
    from srfrd_b200.utils import df_data_partition, interactions_from_df
    df = pd.DataFrame(_fx()["frame"])
    tr, te, un, inum = df_data_partition(df)
    d = interactions_from_df(df)
    assert d.usernum == un and d.itemnum == inum
    for u in range(1, un + 1):
        a, b = d.offsets[u - 1], d.offsets[u]
        assert d.items[a:b].tolist() == tr["item_ids"].get(u, [])
        assert d.labels[a:b].tolist() == tr["review_ids"].get(u, [])
        exp = te["item_ids"].get(u, [])
        assert int(d.test_item[u - 1]) == (exp[0] if exp else 0)
    ref = d.to_reference_dataset()
    assert ref[0]["item_ids"][7] == tr["item_ids"][7]


def test_label_breakdown_restates_reference_aggregation():
    from srfrd_b200.utils import label_breakdown
    rng = np.random.default_rng(3)
    ranks = rng.integers(0, 40, 200)
    labels = rng.integers(0, 4, 200)
    got = label_breakdown(ranks, labels)
    exp = {}
    for lab in sorted(set(labels.tolist())):                 # utils.py:707-752 written out
        rs = ranks[labels == lab]
        ht = sum(1 for r in rs if r < 10)
        nd = sum(1 / np.log2(r + 2) for r in rs if r < 10)
        exp[lab] = [ht / len(rs), nd / len(rs), len(rs)]
    assert list(got) == list(exp)
    for k in exp:
        np.testing.assert_allclose(got[k], exp[k], rtol=1e-12)

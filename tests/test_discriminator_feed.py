"""Discriminator inference feed: a tiny randomly initialised BertForSequenceClassification stands in for the
user's fine-tuned checkpoint (HF transformers is the library the reference calls, data/userDiscriminator.py:10,108).
The batched probabilities reproduce the reference's batch-1 validate() labels, and flow through interactions_from_df into
the CSR the device sampler / loss policies read."""
import numpy as np
import pytest
import torch


def _tiny_bert():
    tr = pytest.importorskip("transformers")
    cfg = tr.BertConfig(vocab_size=200, hidden_size=32, num_hidden_layers=2, num_attention_heads=2, intermediate_size=64,
                        max_position_embeddings=64, num_labels=2)
    torch.manual_seed(0)
    return tr.BertForSequenceClassification(cfg).eval()


def test_probabilities_reproduce_the_reference_validate_labels_and_reach_the_csr():
    import pandas as pd
    from srfrd_b200 import discriminator as D
    from srfrd_b200.utils import interactions_from_df
    from srfrd_b200.trainer import discriminator_weights
    model = _tiny_bert()
    g = torch.Generator().manual_seed(1)
    n, max_len = 60, 24
    ids = torch.randint(1, 200, (n, max_len), generator=g)
    # validate(), data/userDiscriminator.py:57-75, restated: batch size 1, model(ids) without a mask, argmax of the softmax
    ref = []
    with torch.no_grad():
        for i in range(n):
            logits = model(ids[i:i + 1]).logits
            ref.append(int(torch.argmax(torch.nn.functional.softmax(logits, dim=1), dim=1)))
    ref_labels = ["fake" if r == 0 else "real" for r in ref]                     # :117-122
    p = D.fake_probabilities(model, ids, batch_size=16)
    assert p.shape == (n,) and p.dtype == np.float32 and (p >= 0).all() and (p <= 1).all()
    assert D.hard_labels(p) == ref_labels
    # frame -> CSR: the probabilities follow their interactions through the leave-one-out split
    rng = np.random.default_rng(3)
    users = np.repeat(np.arange(1, 13), 5)
    df = pd.DataFrame(dict(user_id=users, item_id=rng.integers(1, 40, n), time=np.tile(np.arange(5), 12)))
    df, pf = D.annotate_frame(df, p)
    assert list(df["fake_review"]) == ref_labels
    data = interactions_from_df(df, p_fake=pf)
    assert data.p_fake.shape == data.items.shape
    # user 1's train interactions are its first 4 rows of the frame
    np.testing.assert_allclose(data.p_fake[data.offsets[0]:data.offsets[1]], pf[:4])
    assert list(data.labels[data.offsets[0]:data.offsets[1]]) == [1 if l == "fake" else 2 for l in ref_labels[:4]]
    # policy 'mask' on the probabilities == the reference's hard label (prs == 2 <=> real)
    pos = torch.ones(1, 4, dtype=torch.long)
    w = discriminator_weights(pos, torch.from_numpy(pf[:4])[None], "mask")
    assert w[0].tolist() == [0.0 if l == "fake" else 1.0 for l in ref_labels[:4]] or np.any(np.isclose(pf[:4], 0.5))


def test_encode_reviews_follows_the_reference_recipe():
    tr = pytest.importorskip("transformers")
    from srfrd_b200 import discriminator as D
    import os, tempfile
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "good", "bad", "product", "fake", "review", "very"]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "vocab.txt")
        open(path, "w").write("\n".join(vocab))
        tok = tr.BertTokenizer(path)
        ids, mask = D.encode_reviews(tok, ["very   good\n product", 12345, "bad " * 40], max_len=16)
    assert ids.shape == (3, 16) and mask.shape == (3, 16) and ids.dtype == torch.long
    assert ids[0, :5].tolist() == [2, 10, 5, 7, 3] and mask[0].sum() == 5          # [CLS] very good product [SEP], padded
    assert mask[2].sum() == 16 and ids[2, -1] == 3                                 # truncated to max_len, ends with [SEP]

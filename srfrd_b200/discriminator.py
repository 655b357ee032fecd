"""Discriminator inference feed: BERT fake-review probabilities -> the per-interaction ``p_fake`` tensor
that row L's loss weights are computed from.

The reference runs a fine-tuned ``BertForSequenceClassification`` over every review with batch size 1, takes the ARGMAX and
writes the strings 'fake' / 'real' into a CSV column (data/userDiscriminator.py:57-75, :117-124; label 0 = fake,
data/trainDiscriminator.py:26-29); the recommender then only sees the hard label.  The hot path here takes the
PROBABILITY (policy 'soft': w = 1 - p_fake; policy 'mask': w = 1[p_fake < 0.5], which is the reference's hard label), so
this module keeps the softmax instead of discarding it.  The transformer itself is a library call (HF ``transformers``, the
checkpoint is the user's): nothing of it is re-implemented -- the arithmetic of this path that the reference owns is the
tokenisation recipe, ``softmax(logits)[:, 0]`` and the label mapping, restated below.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

__all__ = ["encode_reviews", "fake_probabilities", "hard_labels", "annotate_frame"]


def encode_reviews(tokenizer, reviews: Iterable, max_len: int = 512) -> Tuple[torch.Tensor, torch.Tensor]:
    """AmazonReviewDataset.__getitem__ (data/userDiscriminator.py:36-54), batched: str(), whitespace normalised,
    padded / truncated to ``max_len``.  Returns (input_ids, attention_mask), int64 (n, max_len)."""
    texts = [" ".join(str(r).split()) for r in reviews]
    enc = tokenizer(texts, padding="max_length", max_length=max_len, truncation=True, return_tensors="pt")
    return enc["input_ids"].to(torch.long), enc["attention_mask"].to(torch.long)


@torch.no_grad()
def fake_probabilities(model, input_ids: torch.Tensor, batch_size: int = 256, device=None,
                       attention_mask: Optional[torch.Tensor] = None) -> np.ndarray:
    """p_fake per review = softmax(model(ids).logits, dim=1)[:, 0]  (validate(), data/userDiscriminator.py:57-75: the
    reference calls ``model(ids)`` WITHOUT the attention mask and keeps only the argmax; pass ``attention_mask`` to deviate).
    Batched (the reference uses batch size 1, :82); returns float32 (n,)."""
    model.eval()
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    out: List[torch.Tensor] = []
    for s in range(0, input_ids.shape[0], batch_size):
        ids = input_ids[s:s + batch_size].to(dev, dtype=torch.long)
        kw = {}
        if attention_mask is not None:
            kw["attention_mask"] = attention_mask[s:s + batch_size].to(dev, dtype=torch.long)
        logits = model(ids, **kw).logits
        out.append(torch.softmax(logits.float(), dim=1)[:, 0].cpu())
    return torch.cat(out).numpy().astype(np.float32) if out else np.zeros(0, np.float32)


def hard_labels(p_fake: Sequence[float], p_real: Optional[Sequence[float]] = None) -> List[str]:
    """The reference's CSV column (data/userDiscriminator.py:117-122): 'fake' iff argmax == 0.  With two classes that is
    p_fake >= p_real (argmax returns the first maximum), i.e. p_fake >= 0.5 when p_real = 1 - p_fake."""
    pf = np.asarray(p_fake, np.float64)
    pr = 1.0 - pf if p_real is None else np.asarray(p_real, np.float64)
    return ["fake" if a >= b else "real" for a, b in zip(pf, pr)]


def annotate_frame(df, p_fake: np.ndarray):
    """Write the reference's ``fake_review`` column from the probabilities and return (df, p_fake) ready for
    ``srfrd_b200.utils.interactions_from_df(df, p_fake=p_fake)`` -> CSR -> DeviceSampler / FusedTrainer policies."""
    p_fake = np.asarray(p_fake, np.float32)
    if len(p_fake) != len(df):
        raise ValueError(f"{len(p_fake)} probabilities for {len(df)} interactions")
    df = df.copy()
    df["fake_review"] = hard_labels(p_fake)
    return df, p_fake

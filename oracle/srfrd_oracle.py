"""CPU oracle for the SRFRD hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``srfrd_b200`` never imports it and has no CPU fallback.

What it is: a plain fp32 PyTorch/numpy *functional* restatement of the arithmetic
the reference performs on its hot path (embedding gather -> causal self-attention
encoder -> pos/neg scoring + BCE -> catalogue scoring + top-k), written from the
reference's source with every function citing the ``file:line`` it follows
(paths relative to the upstream reference checkout).

Third-party arithmetic: the reference's encoder math lives in PyTorch
(``nn.MultiheadAttention`` need_weights branch, ``nn.LayerNorm``, ``nn.Conv1d``,
``BCEWithLogitsLoss``, ``optim.Adam``); the reference pins no version, so the
oracle is anchored on the installed torch 2.11.0.  The restatement below spells
out that published algorithm with matmul/softmax so it can be compared operand
by operand with the CUDA kernels.

Parity pinning: the reference ships no tests or golden vectors.  The oracle is
pinned against outputs of the *reference itself*, generated in the authoring
container by importing ``SRFR_model.py`` / ``model.py`` unmodified
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``); ``tests/test_oracle.py``
checks restatement == golden for all six model classes.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-8  # SRFR_model.py:77,80,86


# --------------------------------------------------------------------------------------
# model-kind helpers
# --------------------------------------------------------------------------------------
KINDS = ("SRFR", "SRFRN", "SRFU_B", "SRFU_F", "SRFU_R", "SASRec")


def _emb_keys(kind: str) -> Tuple[str, str]:
    """state_dict names of the item / positional tables (SRFR_model.py:10-12 vs :591-592)."""
    if kind == "SASRec":
        return "item_emb.weight", "pos_emb.weight"
    return "embedding_layer.item_embed.weight", "embedding_layer.pos_embed.weight"


def init_state_dict(kind: str, item_number: int, max_len: int, D: int, F: int = 0, n_labels: int = 0, num_blocks: int = 2,
                    seed: int = 0) -> Dict[str, Tensor]:
    """A state_dict with the reference's parameter names and shapes (dumped from the live modules) and the
    trainer's initialisation: xavier_normal_ on every >= 2-D parameter (trainer.py:364-369, which also overwrites the
    padding rows), LayerNorm weight 1 / bias 0, other biases 0.  Lets the CPU baseline run without the product package."""
    g = torch.Generator().manual_seed(seed)
    H = D + F if kind in ("SRFR", "SRFRN") else D
    shapes = []
    if kind == "SASRec":
        shapes += [("item_emb.weight", (item_number + 1, D)), ("pos_emb.weight", (max_len, D))]
    elif kind in ("SRFR", "SRFRN"):
        shapes += [("embedding_layer.item_embed.weight", (item_number + 1, D)), ("embedding_layer.fake_embed.weight", (3, F)),
                   ("embedding_layer.pos_embed.weight", (max_len, D))]
    else:
        shapes += [("embedding_layer.item_embed.weight", (item_number + 1, D)),
                   ("embedding_layer.user_label_embed.weight", (n_labels, D)), ("embedding_layer.pos_embed.weight", (max_len, D))]
    for i in range(num_blocks):
        shapes += [(f"attention_layernorms.{i}.weight", (H,)), (f"attention_layernorms.{i}.bias", (H,)),
                   (f"attention_layers.{i}.in_proj_weight", (3 * H, H)), (f"attention_layers.{i}.in_proj_bias", (3 * H,)),
                   (f"attention_layers.{i}.out_proj.weight", (H, H)), (f"attention_layers.{i}.out_proj.bias", (H,)),
                   (f"forward_layernorms.{i}.weight", (H,)), (f"forward_layernorms.{i}.bias", (H,)),
                   (f"forward_layers.{i}.conv1.weight", (H, H, 1)), (f"forward_layers.{i}.conv1.bias", (H,)),
                   (f"forward_layers.{i}.conv2.weight", (H, H, 1)), (f"forward_layers.{i}.conv2.bias", (H,))]
    if kind == "SRFR":
        shapes += [("last_conv.weight", (D, H, 1)), ("last_conv.bias", (D,))]
    dout = H if kind == "SRFRN" else D
    shapes += [("last_layernorm.weight", (dout,)), ("last_layernorm.bias", (dout,))]
    sd = {}
    for name, shape in shapes:
        t = torch.zeros(shape)
        if len(shape) >= 2:
            torch.nn.init.xavier_normal_(t, generator=g)
        elif "layernorm" in name and name.endswith("weight"):
            t.fill_(1.0)
        sd[name] = t
    return sd


def num_blocks_of(sd: Dict[str, Tensor]) -> int:
    n = 0
    while f"attention_layers.{n}.in_proj_weight" in sd:
        n += 1
    return n


def srfu_labels(kind: str, fake_ids: Tensor) -> Tensor:
    """SRFU_B/F/R.get_Labels, SRFR_model.py:546-570."""
    nf = torch.count_nonzero(fake_ids == 1, dim=1)
    nr = torch.count_nonzero(fake_ids == 2, dim=1)
    if kind == "SRFU_B":
        return torch.round(torch.sign(nf - nr) * 0.5 + 1.5).int()
    if kind == "SRFU_F":
        return nf
    if kind == "SRFU_R":
        return torch.floor(nf / (nf + nr) * 10).int()
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# A1/A2/A3: embedding gather + positional add (+ fake concat / user-label add) + pad mask
# --------------------------------------------------------------------------------------
def embed(sd: Dict[str, Tensor], kind: str, input_ids: Tensor, fake_ids: Optional[Tensor]) -> Tensor:
    """Pre-LayerNorm encoder input x0 (B, L, H), fp32.

    SRFR/SRFRN: SRFR_Embedding.forward SRFR_model.py:17-34 then pad mask :98-99.
    SRFU_*:     SRFU_Embedding.forward :411-424 then :485-487.
    SASRec:     log2feats :620-628 (scaled by sqrt(d), eval-mode dropout = identity).
    """
    ik, pk = _emb_keys(kind)
    E, P = sd[ik], sd[pk]
    B, L = input_ids.shape
    x = F.embedding(input_ids.long(), E, padding_idx=0)      # nn.Embedding(..., padding_idx=0), SRFR_model.py:10
    if kind == "SASRec":
        x = x * (E.shape[1] ** 0.5)
    x = x + P[torch.arange(L, device=P.device)].unsqueeze(0)
    if kind in ("SRFR", "SRFRN"):
        Fe = sd["embedding_layer.fake_embed.weight"]
        if fake_ids is None:
            fake_ids = torch.zeros(B, L, dtype=torch.long, device=E.device)
        x = torch.cat([x, F.embedding(fake_ids.long(), Fe, padding_idx=0)], dim=2)       # SRFR_model.py:11
    elif kind.startswith("SRFU"):
        Ul = sd["embedding_layer.user_label_embed.weight"]
        lab = srfu_labels(kind, fake_ids).long().view(B, 1)
        x = x + Ul[lab]
    x = x * (input_ids != 0).unsqueeze(-1)
    return x


# --------------------------------------------------------------------------------------
# A4: one encoder block (non-standard wiring, SRFR_model.py:109-121)
# --------------------------------------------------------------------------------------
def attention(q: Tensor, k: Tensor, v: Tensor, num_heads: int) -> Tensor:
    """Causal multi-head attention, no key-padding mask (SRFR_model.py:103,112-113).

    Follows F.multi_head_attention_forward's need_weights branch: q is scaled by
    head_dim**-0.5 BEFORE q k^T, -inf above the diagonal, softmax, (dropout omitted:
    parity runs use p=0 / eval), then @ v.
    q, k, v: (B, L, H) -> (B, L, H)
    """
    B, L, H = q.shape
    hd = H // num_heads
    qh = q.view(B, L, num_heads, hd).transpose(1, 2) * (1.0 / math.sqrt(hd))
    kh = k.view(B, L, num_heads, hd).transpose(1, 2)
    vh = v.view(B, L, num_heads, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    causal = torch.ones(L, L, dtype=torch.bool, device=s.device).tril()
    s = s.masked_fill(~causal, float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = a @ vh
    return o.transpose(1, 2).reshape(B, L, H)


def encoder_block(sd: Dict[str, Tensor], i: int, x: Tensor, mask: Tensor, num_heads: int,
                  trace: Optional[dict] = None) -> Tensor:
    """SRFR_model.py:109-121: Q = LN1(x); q from Q, k/v from UN-normalised x; residual adds Q;
    LN2 output is the FFN residual stream (PointWiseFeedForward :47-51); re-mask pads."""
    H = x.shape[-1]
    W = sd[f"attention_layers.{i}.in_proj_weight"]
    b = sd[f"attention_layers.{i}.in_proj_bias"]
    Q = F.layer_norm(x, (H,), sd[f"attention_layernorms.{i}.weight"], sd[f"attention_layernorms.{i}.bias"], LN_EPS)
    q = Q @ W[:H].T + b[:H]
    k = x @ W[H:2 * H].T + b[H:2 * H]
    v = x @ W[2 * H:].T + b[2 * H:]
    o = attention(q, k, v, num_heads)
    o = o @ sd[f"attention_layers.{i}.out_proj.weight"].T + sd[f"attention_layers.{i}.out_proj.bias"]
    r = Q + o
    y = F.layer_norm(r, (H,), sd[f"forward_layernorms.{i}.weight"], sd[f"forward_layernorms.{i}.bias"], LN_EPS)
    W1 = sd[f"forward_layers.{i}.conv1.weight"].squeeze(-1)
    W2 = sd[f"forward_layers.{i}.conv2.weight"].squeeze(-1)
    h1 = torch.relu(y @ W1.T + sd[f"forward_layers.{i}.conv1.bias"])
    z = h1 @ W2.T + sd[f"forward_layers.{i}.conv2.bias"] + y
    out = z * mask
    if trace is not None:
        trace[f"Q{i}"], trace[f"q{i}"], trace[f"k{i}"], trace[f"v{i}"] = Q, q, k, v
        trace[f"r{i}"], trace[f"y{i}"], trace[f"h1_{i}"], trace[f"x{i + 1}"] = r, y, h1, out
    return out


def encode(sd: Dict[str, Tensor], kind: str, input_ids: Tensor, fake_ids: Optional[Tensor],
           num_heads: int = 1, trace: Optional[dict] = None) -> Tensor:
    """Full encoder: hidden_state (B, L, D or H).  SRFR.forward SRFR_model.py:92-124;
    SRFRN :190-221; SRFU :475-510; SASRec.log2feats :620-649."""
    x = embed(sd, kind, input_ids, fake_ids)
    if trace is not None:
        trace["x0"] = x
    mask = (input_ids != 0).unsqueeze(-1)
    for i in range(num_blocks_of(sd)):
        x = encoder_block(sd, i, x, mask, num_heads, trace)
    if kind == "SRFR":  # last_conv (H -> D) then last_layernorm, :123-124
        Wc = sd["last_conv.weight"].squeeze(-1)
        x = x @ Wc.T + sd["last_conv.bias"]
    D = x.shape[-1]
    return F.layer_norm(x, (D,), sd["last_layernorm.weight"], sd["last_layernorm.bias"], LN_EPS)


# --------------------------------------------------------------------------------------
# A6: pos / neg logits
# --------------------------------------------------------------------------------------
def target_rows(sd: Dict[str, Tensor], kind: str, ids: Tensor, fake_ids: Optional[Tensor]) -> Tensor:
    """Rows the hidden state is dotted with: E[id] (SRFR :129-136, SRFU :516-525, SASRec :657-658)
    or E[id] || Fe[fake_id] for SRFRN (:225,:231)."""
    E = sd[_emb_keys(kind)[0]]
    rows = F.embedding(ids.long(), E, padding_idx=0)         # the same nn.Embedding module is called (SRFR_model.py:129)
    if kind == "SRFRN":
        rows = torch.cat([rows, F.embedding(fake_ids.long(), sd["embedding_layer.fake_embed.weight"], padding_idx=0)], dim=-1)
    return rows


def forward(sd, kind, input_ids, fake_ids, pos=None, prs=None, neg=None, nrs=None, num_heads=1):
    """(hidden, pos_logits, neg_logits) as the reference forward returns them."""
    h = encode(sd, kind, input_ids, fake_ids, num_heads)
    zp = (h * target_rows(sd, kind, pos, prs)).sum(-1) if pos is not None else None
    zn = (h * target_rows(sd, kind, neg, nrs)).sum(-1) if neg is not None else None
    return h, zp, zn


# --------------------------------------------------------------------------------------
# A7 + row L: masked / discriminator-weighted BCE
# --------------------------------------------------------------------------------------
def reference_loss(pos_logits: Tensor, neg_logits: Tensor, pos: Tensor) -> Tensor:
    """trainer.py:36-38 verbatim semantics: mean BCEWithLogits over positions with pos != 0
    (l2_emb term :39 is 0.0 by default and added separately by callers)."""
    idx = torch.where(pos != 0)
    crit = torch.nn.BCEWithLogitsLoss()
    loss = crit(pos_logits[idx], torch.ones_like(pos_logits[idx]))
    loss = loss + crit(neg_logits[idx], torch.zeros_like(neg_logits[idx]))
    return loss


def discriminator_weights(pos: Tensor, p_fake: Optional[Tensor], policy: str) -> Tensor:
    """Row L (extension; not in the reference): w = 1[pos != 0] * g(p_fake).
    'none' g=1 (== reference), 'mask' g = 1[p_fake < 0.5], 'soft' g = 1 - p_fake."""
    valid = (pos != 0).float()
    if policy == "none" or p_fake is None:
        return valid
    if policy == "mask":
        return valid * (p_fake < 0.5).float()
    if policy == "soft":
        return valid * (1.0 - p_fake.float())
    raise ValueError(policy)


def weighted_loss(pos_logits: Tensor, neg_logits: Tensor, w_pos: Tensor, w_neg: Optional[Tensor] = None) -> Tensor:
    """sum(w * softplus(-z+)) / sum(w) + sum(w' * softplus(z-)) / sum(w').
    With w = w' = 1[pos != 0] this equals reference_loss (trainer.py:36-38)."""
    if w_neg is None:
        w_neg = w_pos
    lp = (w_pos * F.softplus(-pos_logits)).sum() / w_pos.sum()
    ln = (w_neg * F.softplus(neg_logits)).sum() / w_neg.sum()
    return lp + ln


# --------------------------------------------------------------------------------------
# A9/A10 + full-catalogue extension
# --------------------------------------------------------------------------------------
def predict(sd, kind, input_ids, fake_ids, label, num_heads=1) -> Tensor:
    """predict(): last-position hidden dotted with E[label]  (SRFR_model.py:144-152).
    Returned un-squeezed as (U, I)."""
    h = encode(sd, kind, input_ids, fake_ids, num_heads)[:, -1, :]
    E = sd[_emb_keys(kind)[0]]
    return h @ E[label.long()].T


def topk_stable(scores: np.ndarray, k: int, first_id: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k per row with the tie-break (score desc, item id asc).  torch.topk does not promise
    a tie order, so the oracle uses a stable sort on -score (column order == id order)."""
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    vals = np.take_along_axis(scores, order, axis=1)
    return vals, (order + first_id).astype(np.int64)


def catalogue_topk(feats: Tensor, table: Tensor, k: int = 10, chunk: int = 2048,
                   operand_dtype: Optional[torch.dtype] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Full-catalogue scoring: feats (U, D) against table rows 1..N (row 0 = pad, excluded);
    == predict(..., label=arange(1, N+1)) followed by a stable top-k.
    operand_dtype=torch.bfloat16 rounds BOTH operands to bf16 first (what the tensor-core
    kernel consumes) and accumulates in fp64 so the accumulation order cannot matter."""
    f, t = feats.float(), table[1:].float()
    acc = torch.float32
    if operand_dtype is not None:
        f, t = f.to(operand_dtype).double(), t.to(operand_dtype).double()
        acc = torch.float64
    vals, ids = [], []
    for s in range(0, f.shape[0], chunk):
        sc = (f[s:s + chunk].to(acc) @ t.to(acc).T).numpy()
        v, i = topk_stable(sc, k)
        vals.append(v)
        ids.append(i)
    return np.concatenate(vals), np.concatenate(ids)


def merge_topk(vals: np.ndarray, ids: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Merge per-shard candidate lists (U, S*k) -> (U, k), tie-break (score desc, id asc)."""
    order = np.lexsort((ids, -vals), axis=1)[:, :k]
    return np.take_along_axis(vals, order, 1), np.take_along_axis(ids, order, 1)


def hr_ndcg_from_rank(rank: np.ndarray, k: int = 10) -> Tuple[float, float]:
    """utils.py:595-597: HT += 1, NDCG += 1/log2(rank+2) if rank < 10; means over valid users.
    Returns (NDCG@k, HR@k) in the reference's return order (utils.py:602)."""
    hit = rank < k
    ndcg = np.where(hit, 1.0 / np.log2(rank + 2.0), 0.0)
    n = max(len(rank), 1)
    return float(ndcg.sum() / n), float(hit.sum() / n)


def rank_of_first(scores: np.ndarray) -> np.ndarray:
    """utils.py:589-591: rank of candidate 0 = (-logits).argsort().argsort()[0] per row."""
    return np.argsort(np.argsort(-scores, axis=1, kind="stable"), axis=1, kind="stable")[:, 0]


def full_catalogue_rank(feats: Tensor, table: Tensor, target: np.ndarray) -> np.ndarray:
    """Rank of `target` among items 1..N under (score desc, id asc)."""
    sc = (feats.float() @ table[1:].float().T).numpy()
    tgt = sc[np.arange(len(target)), target - 1]
    better = (sc > tgt[:, None]).sum(1)
    ties_before = ((sc == tgt[:, None]) & (np.arange(1, sc.shape[1] + 1)[None, :] < target[:, None])).sum(1)
    return better + ties_before


# --------------------------------------------------------------------------------------
# A8: one training step (trainer.py:27-41) on top of torch autograd + Adam
# --------------------------------------------------------------------------------------
class OracleTrainer:
    """Holds fp32 parameters (a state_dict cloned into leaf tensors) and runs the reference's
    step: forward -> masked/weighted BCE -> backward -> Adam(lr, betas=(0.9, 0.98)) trainer.py:390.
    The item / fake tables get padding_idx=0 semantics: gradient row 0 is zeroed
    (SRFR_model.py:10-11, nn.Embedding padding_idx)."""

    def __init__(self, sd: Dict[str, Tensor], kind: str, num_heads: int = 1, lr: float = 1e-3,
                 betas=(0.9, 0.98), eps: float = 1e-8):
        self.kind, self.num_heads = kind, num_heads
        self.sd = {k: v.detach().clone().float().requires_grad_(True) for k, v in sd.items()}
        self.opt = torch.optim.Adam(list(self.sd.values()), lr=lr, betas=betas, eps=eps)

    def _zero_pad_rows(self):
        names = [_emb_keys(self.kind)[0]]
        if self.kind in ("SRFR", "SRFRN"):
            names.append("embedding_layer.fake_embed.weight")
        for n in names:
            if self.sd[n].grad is not None:
                self.sd[n].grad[0].zero_()

    def loss(self, batch: dict, w_pos: Optional[Tensor] = None, w_neg: Optional[Tensor] = None) -> Tensor:
        _, zp, zn = forward(self.sd, self.kind, batch["seq"], batch["rsq"], batch["pos"], batch["prs"],
                            batch["neg"], batch["nrs"], self.num_heads)
        if w_pos is None:
            return reference_loss(zp, zn, batch["pos"])
        return weighted_loss(zp, zn, w_pos, w_neg)

    def step(self, batch: dict, w_pos: Optional[Tensor] = None, w_neg: Optional[Tensor] = None) -> float:
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss(batch, w_pos, w_neg)
        loss.backward()
        self._zero_pad_rows()
        self.opt.step()
        return float(loss.detach())

    def grads(self, batch: dict, w_pos=None, w_neg=None) -> Tuple[float, Dict[str, Tensor]]:
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss(batch, w_pos, w_neg)
        loss.backward()
        self._zero_pad_rows()
        return float(loss.detach()), {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v))
                                      for k, v in self.sd.items()}

    def state_dict(self) -> Dict[str, Tensor]:
        return {k: v.detach().clone() for k, v in self.sd.items()}

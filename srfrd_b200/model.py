"""Drop-in for the reference's legacy ``model.py`` SASRec (numpy inputs, ``args`` namespace).

Same constructor ``SASRec(user_num, item_num, args)`` reading ``args.{device, hidden_units, maxlen,
dropout_rate, num_blocks, num_heads}`` (model.py:27-60), ``forward(user_ids, log_seqs, pos_seqs,
neg_seqs) -> (pos_logits, neg_logits)`` (model.py:96-108) and ``predict(user_ids, log_seqs,
item_indices) -> (U, I)`` (model.py:110-120).  The arithmetic is identical to
``SRFR_model.SASRec`` and runs on the same kernels.
"""
from __future__ import annotations

import numpy as np
import torch

from . import SRFR_model as _M


class SASRec(_M.SASRec):
    def __init__(self, user_num, item_num, args):
        super().__init__(item_num, args.maxlen, args.hidden_units, args.dropout_rate, args.num_blocks, args.num_heads,
                         args.device)
        self.user_num = user_num

    def _t(self, a):
        dev = next(self.parameters()).device
        return torch.as_tensor(np.asarray(a)).long().to(dev)

    def log2feats(self, log_seqs):
        return super().forward(None, self._t(log_seqs), None)[0]

    def forward(self, user_ids, log_seqs, pos_seqs, neg_seqs):  # for training
        _, zp, zn = super().forward(None, self._t(log_seqs), None, self._t(pos_seqs), None, self._t(neg_seqs), None)
        return zp, zn

    def predict(self, user_ids, log_seqs, item_indices):  # for inference
        out = super().predict(None, self._t(log_seqs), None, self._t(item_indices))
        U = np.asarray(log_seqs).shape[0]
        return out.reshape(U, -1)

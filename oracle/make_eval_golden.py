"""Golden for the sampled-101 evaluator: run the UNMODIFIED reference ``evaluation()`` (utils.py:544-602, exec'd from
the upstream checkout because utils.py as a whole does not parse, SURVEY.md note 4) with the UNMODIFIED reference model
``SRFR_model.SRFR`` on a seeded synthetic dataset, and record what it computed.

TEST INFRASTRUCTURE; needs /root/reference (authoring container only).  Writes tests/golden/eval101.npz:
  users      evaluated user ids in the reference's loop order
  ranks      the rank the reference computed for each of them (utils.py:591)
  ndcg, hr   what evaluation() returned
  cand_crc   crc32 of the (U, 101) candidate matrix the reference drew (np.random.seed(SEED) + utils.py:576-583);
             the tests re-draw it with the same legacy RandomState stream (restated loop) and check the crc
  cand_head  the first 8 users' candidate rows verbatim
The model weights are oracle.init_state_dict("SRFR", ..., seed=SEED) loaded into the reference module, and the data set
is srfrd_b200.synth.make_interactions(DATA_SEED, ...) -- both deterministic, so the tests rebuild them instead of
storing ~1 MB of parameters (a crc of the weights is stored to detect RNG drift).
"""
import contextlib
import copy
import io
import os
import random
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("SRFRD_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "eval101.npz")

SEED, DATA_SEED = 4242, 77
USERS, ITEMS, L, D, F, NB = 4000, 3000, 50, 64, 16, 2


def crc_state(sd):
    c = 0
    for k in sorted(sd):
        c = zlib.crc32(sd[k].detach().cpu().numpy().tobytes(), c)
    return c


def main():
    from oracle import srfrd_oracle as O
    from srfrd_b200 import synth
    sys.path.insert(0, REF)
    import SRFR_model as SR
    src = open(os.path.join(REF, "utils.py")).read().splitlines()
    ns = {"np": np, "torch": torch, "copy": copy, "random": random, "sys": sys}
    exec("\n".join(src[543:602]), ns)                     # def evaluation(model, dataset, maxlen, device): ... return
    data = synth.make_interactions(DATA_SEED, USERS, ITEMS, 3, 8.0, L)
    dataset = data.to_reference_dataset()
    sd = O.init_state_dict("SRFR", ITEMS, L, D, F, 0, NB, seed=SEED)
    model = SR.SRFR(ITEMS, L, D, F, 0.0, NB, 1, "cpu")
    model.load_state_dict(sd)
    model.eval()
    rec_cand, rec_rank, rec_user = [], [], []
    real_predict = model.predict

    def spy(u, seq, rsq, label):
        out = real_predict(u, seq, rsq, label)
        rec_cand.append(label.numpy().copy())
        rec_rank.append(int((-out).argsort().argsort()[0].item()))
        return out

    model.predict = spy
    # the reference wraps the user id in torch.LongTensor(u) (an uninitialised tensor of u elements): harmless, unused
    np.random.seed(SEED)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        ndcg, hr = ns["evaluation"](model, dataset, L, "cpu")
    users = [u for u in range(1, USERS + 1) if len(dataset[0]["item_ids"][u]) >= 1 and len(dataset[1]["item_ids"][u]) >= 1]
    assert len(users) == len(rec_rank)
    cand = np.stack(rec_cand).astype(np.int64)
    np.savez_compressed(OUT, users=np.asarray(users, np.int32), ranks=np.asarray(rec_rank, np.int16), ndcg=float(ndcg),
                        hr=float(hr), cand_crc=np.int64(zlib.crc32(cand.tobytes())), cand_head=cand[:8],
                        state_crc=np.int64(crc_state(sd)),
                        meta=np.array([SEED, DATA_SEED, USERS, ITEMS, L, D, F, NB], np.int64))
    print(f"wrote {OUT}: {len(users)} users, NDCG@10 {ndcg:.5f} HR@10 {hr:.5f}")


if __name__ == "__main__":
    main()

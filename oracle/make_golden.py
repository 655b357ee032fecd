"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

TEST INFRASTRUCTURE.  Runs only in the authoring container (needs the upstream checkout at
/root/reference, which does not exist on the GPU box); the resulting small fixtures are
committed so GPU-side tests never read the reference at run time.

    python oracle/make_golden.py            # rewrites tests/golden/

Each fixture holds, for one model class: the state_dict, one seeded batch in the sampler's
layout (utils.py:21-65), and what the reference computes from them:
  hidden / pos_logits / neg_logits      <- model(...)                SRFR_model.py:92-142
  loss                                  <- trainer.py:36-38
  grad.<param>                          <- loss.backward()           trainer.py:40
  loss_steps, after.<param>             <- 3 x (zero_grad, backward, Adam(1e-3,(0.9,0.98)).step())
  predict_all                           <- model.predict(None, seq, rsq, arange(1, N+1))
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SRFRD_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import SRFR_model  # noqa
    import model as ref_model  # noqa
    return SRFR_model, ref_model


def init_like_trainer(model, gen):
    """trainer.py:364-369: xavier_normal_ on every >=2-D parameter (this also overwrites the zero
    padding rows).  1-D parameters (LN weight/bias, biases) are additionally jittered so the
    fixtures exercise them; a trained checkpoint would have non-trivial values there too."""
    for _, p in model.named_parameters():
        if p.dim() >= 2:
            torch.nn.init.xavier_normal_(p.data, generator=gen)
        else:
            p.data.add_(0.1 * torch.randn(p.shape, generator=gen))


def make_batch(rng, B, L, N):
    seq = np.zeros((B, L), np.int64); pos = np.zeros_like(seq); neg = np.zeros_like(seq)
    rsq = np.zeros_like(seq); prs = np.zeros_like(seq); nrs = np.zeros_like(seq)
    for b in range(B):
        n = int(rng.integers(2, L + 3))                      # some users longer than L (truncation)
        items = rng.permutation(N)[:n] + 1
        labs = rng.integers(1, 3, n)
        m = min(n - 1, L)
        seq[b, L - m:] = items[n - 1 - m:n - 1]; pos[b, L - m:] = items[n - m:n]
        rsq[b, L - m:] = labs[n - 1 - m:n - 1]; prs[b, L - m:] = labs[n - m:n]
        pool = np.setdiff1d(np.arange(1, N + 1), items)
        neg[b, L - m:] = rng.choice(pool, m); nrs[b, L - m:] = 1
    if B > 1:
        seq[1, :] = 0; pos[1, :] = 0; neg[1, :] = 0; rsq[1, :] = 0; prs[1, :] = 0; nrs[1, :] = 0
        seq[1, -1] = 3; pos[1, -1] = 4; neg[1, -1] = 5; rsq[1, -1] = 2; prs[1, -1] = 1; nrs[1, -1] = 1
    return dict(seq=seq, rsq=rsq, pos=pos, prs=prs, neg=neg, nrs=nrs)


def build(kind, SR, N, L, D, Fw, nb, heads):
    if kind == "SRFR":
        return SR.SRFR(N, L, D, Fw, 0.0, nb, heads, "cpu")
    if kind == "SRFRN":
        return SR.SRFRN(N, L, D, Fw, 0.0, nb, heads, "cpu")
    if kind == "SRFU_B":
        return SR.SRFU_B(N, L, D, 3, 0.0, nb, heads, "cpu")
    if kind == "SRFU_F":
        return SR.SRFU_F(N, L, D, L + 1, 0.0, nb, heads, "cpu")
    if kind == "SRFU_R":
        return SR.SRFU_R(N, L, D, 11, 0.0, nb, heads, "cpu")
    if kind == "SASRec":
        return SR.SASRec(N, L, D, 0.0, nb, heads, "cpu")
    raise ValueError(kind)


def one_fixture(kind, SR, seed, N=60, L=12, D=16, Fw=16, nb=2, heads=1, B=6):
    torch.manual_seed(seed)                  # the constructors' default init (conv biases ...) draws from the GLOBAL stream
    gen = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    model = build(kind, SR, N, L, D, Fw, nb, heads)
    init_like_trainer(model, gen)
    model.train()                                             # dropout_rate = 0.0 -> deterministic
    batch = make_batch(rng, B, L, N)
    if kind == "SRFU_R":                                      # 0/0 -> NaN label for all-pad rows (:566)
        batch["rsq"][batch["rsq"].sum(1) == 0, -1] = 2
    t = {k: torch.from_numpy(v) for k, v in batch.items()}
    out = {"meta": np.array([N, L, D, Fw, nb, heads, B], np.int64)}
    for k, v in model.state_dict().items():
        out["param." + k] = v.numpy().copy()
    for k, v in batch.items():
        out["in." + k] = v

    def fwd_loss():
        h, zp, zn = model(None, t["seq"], t["rsq"], t["pos"], t["prs"], t["neg"], t["nrs"])
        idx = torch.where(t["pos"] != 0)                      # trainer.py:36-38
        crit = torch.nn.BCEWithLogitsLoss()
        loss = crit(zp[idx], torch.ones_like(zp[idx])) + crit(zn[idx], torch.zeros_like(zn[idx]))
        return h, zp, zn, loss

    h, zp, zn, loss = fwd_loss()
    out["hidden"], out["pos_logits"], out["neg_logits"] = h.detach().numpy(), zp.detach().numpy(), zn.detach().numpy()
    out["loss"] = np.float32(loss.item())
    model.zero_grad()
    loss.backward()
    for k, p in model.named_parameters():
        out["grad." + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    with torch.no_grad():
        model.eval()
        lab = torch.arange(1, N + 1)
        if kind == "SRFRN":                                   # SRFRN.predict is only valid for U=1 (:241-259)
            pa = torch.stack([model.predict(None, t["seq"][b:b + 1], t["rsq"][b:b + 1], lab) for b in range(B)])
        else:
            pa = model.predict(None, t["seq"], t["rsq"], lab)
        out["predict_all"] = pa.numpy().copy()
        model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.98))   # trainer.py:390
    losses = []
    for _ in range(3):
        _, _, _, loss = fwd_loss()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    out["loss_steps"] = np.array(losses, np.float32)
    for k, v in model.state_dict().items():
        out["after." + k] = v.numpy().copy()
    return out


def legacy_sasrec_fixture(RM, seed, N=60, L=12, D=16, nb=2, B=6):
    """model.py SASRec (numpy inputs, args namespace)."""
    import types
    args = types.SimpleNamespace(device="cpu", hidden_units=D, maxlen=L, dropout_rate=0.0, num_blocks=nb, num_heads=1)
    torch.manual_seed(seed)                  # the constructors' default init (conv biases ...) draws from the GLOBAL stream
    gen = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    model = RM.SASRec(10, N, args)
    init_like_trainer(model, gen)
    batch = make_batch(rng, B, L, N)
    out = {"meta": np.array([N, L, D, 0, nb, 1, B], np.int64)}
    for k, v in model.state_dict().items():
        out["param." + k] = v.numpy().copy()
    for k, v in batch.items():
        out["in." + k] = v
    zp, zn = model(None, batch["seq"], batch["pos"], batch["neg"])
    out["pos_logits"], out["neg_logits"] = zp.detach().numpy(), zn.detach().numpy()
    with torch.no_grad():
        out["predict_all"] = model.predict(None, batch["seq"], np.arange(1, N + 1)).numpy().copy()
    return out


def main():
    SR, RM = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    for i, kind in enumerate(["SRFR", "SRFRN", "SRFU_B", "SRFU_F", "SRFU_R", "SASRec"]):
        fx = one_fixture(kind, SR, 100 + i)
        np.savez_compressed(os.path.join(OUT, f"{kind}.npz"), **fx)
        print(kind, "loss", fx["loss"], "steps", fx["loss_steps"])
    fx = one_fixture("SRFR", SR, 200, N=90, L=10, D=16, Fw=16, nb=1, heads=2, B=5)
    np.savez_compressed(os.path.join(OUT, "SRFR_heads2.npz"), **fx)
    fx = one_fixture("SASRec", SR, 201, N=40, L=16, D=32, nb=3, heads=4, B=4)
    np.savez_compressed(os.path.join(OUT, "SASRec_heads4.npz"), **fx)
    fx = legacy_sasrec_fixture(RM, 300)
    np.savez_compressed(os.path.join(OUT, "legacy_SASRec.npz"), **fx)
    ragged_fixtures(SR)
    print("wrote", OUT)


def ragged_fixtures(SR):
    """Widths that are not multiples of 16: the author's own run (45 + 5, trainer.py:129-130) and the constructor
    defaults (50 + 10, SRFR_model.py:54-63)."""
    for name, kind, kw in (("SRFR_w50", "SRFR", dict(D=45, Fw=5)), ("SRFRN_w60", "SRFRN", dict(D=50, Fw=10)),
                           ("SASRec_w50", "SASRec", dict(D=50)), ("SRFU_B_w50", "SRFU_B", dict(D=50))):
        fx = one_fixture(kind, SR, 400 + len(name), N=70, L=14, nb=2, heads=1, B=5, **kw)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **fx)
        print(name, "loss", fx["loss"], "steps", fx["loss_steps"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ragged":      # only add the ragged-width fixtures
        torch.set_num_threads(1)
        ragged_fixtures(load_reference()[0])
    else:
        main()

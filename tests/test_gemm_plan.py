"""Host logic of srfrd_gemm_tn (tile / pipeline plan) checked on the CPU through srfrd_gemm_tn_plan, and the real plans of
the benchmark shapes run through the pipeline-protocol model (tests/test_pipeline_protocol.py)."""
import ctypes
import itertools

import pytest

from tests.test_pipeline_protocol import RING, Sim

SMEM_MAX = 227 * 1024
KEYS = ("block_n", "n_tiles", "stages", "kgroup", "b_resident", "nacc", "nbuf", "buf_blocks", "smem", "two_issuers", "kblocks")


@pytest.fixture(scope="module")
def plan():
    from srfrd_b200 import _lib
    lib = _lib.load()

    def f(M, N, K, aux=False, bf16_out=True, ln=False):
        out = (ctypes.c_int * 11)()
        rc = lib.srfrd_gemm_tn_plan(M, N, K, int(aux), int(bf16_out), int(ln), out)
        if rc != 0:
            raise RuntimeError(lib.srfrd_last_error().decode())
        return dict(zip(KEYS, out))
    return f


SHAPES = [(N, K) for N, K in itertools.product((16, 48, 64, 80, 96, 128, 160, 192, 240, 272, 544), (16, 64, 80, 96, 128, 160, 272))]


@pytest.mark.parametrize("aux,bf16_out,ln", [(False, True, False), (True, True, False), (True, True, True), (False, False, False)])
def test_plan_invariants(plan, aux, bf16_out, ln):
    for N, K in SHAPES:
        if ln and N > 128:                 # the launcher accepts the fused LayerNorm only up to 128 columns
            continue
        p = plan(204800, N, K, aux, bf16_out, ln)
        assert p["smem"] <= SMEM_MAX, (N, K, p)
        assert p["stages"] >= 2 and p["stages"] <= 6
        assert p["kblocks"] == (K + 63) // 64 and p["kgroup"] in (1, p["kblocks"])
        assert p["block_n"] % 16 == 0 and p["block_n"] * p["n_tiles"] >= N and p["block_n"] <= 192
        assert (p["nacc"] == 4) == (p["block_n"] <= 128)                 # 4 x 128 or 2 x 256 TMEM columns
        assert p["nacc"] * (128 if p["nacc"] == 4 else 256) <= 512
        if p["b_resident"]:
            assert p["n_tiles"] == 1
        if p["kgroup"] > 1:
            assert p["b_resident"]                                       # whole-K stages only with the weights resident
        assert p["nbuf"] in (2, 4) and (p["nbuf"] == 2 or (aux and not ln and p["n_tiles"] == 1))
        assert p["two_issuers"] == int(p["kgroup"] >= p["kblocks"] and p["stages"] % 2 == 0)
        # tile-id ring: producer <= ceil(stages * kgroup / kblocks) tiles ahead of the MMAs, MMAs <= nacc ahead of the
        # epilogues, plus the two end markers
        ahead = -(-p["stages"] * p["kgroup"] // p["kblocks"])
        assert ahead + p["nacc"] + 2 < RING, (N, K, p)


def test_plan_of_the_benchmark_shapes(plan):
    """C2 (H = 80): weights resident, whole-K stages, two issuers; residual tiles get two buffers per set"""
    p = plan(204800, 80, 80)
    assert (p["b_resident"], p["kgroup"], p["stages"], p["two_issuers"], p["nacc"], p["nbuf"]) == (1, 2, 4, 1, 4, 2)
    p = plan(204800, 80, 80, aux=True)
    assert (p["kgroup"], p["stages"], p["nbuf"], p["two_issuers"]) == (2, 2, 4, 1)
    p = plan(204800, 80, 80, aux=True, ln=True)
    assert (p["kgroup"], p["stages"], p["nbuf"]) == (2, 2, 2)
    p = plan(204800, 272, 272, aux=True)                                  # C4: two column tiles, weights through the ring
    assert (p["n_tiles"], p["b_resident"], p["kgroup"], p["nacc"], p["nbuf"]) == (2, 0, 1, 2, 2)


def test_plan_rejects_what_the_kernel_cannot_do(plan):
    with pytest.raises(RuntimeError, match="one column tile"):
        plan(1000, 272, 80, aux=True, ln=True)


@pytest.mark.parametrize("N,K,aux,ln", [(80, 80, False, False), (80, 80, True, False), (80, 80, True, True), (160, 80, False, False),
                                        (80, 160, False, False), (64, 80, False, False), (272, 272, True, False),
                                        (128, 128, True, True), (192, 64, True, False)])
def test_real_plans_pass_the_protocol_model(plan, N, K, aux, ln):
    p = plan(204800, N, K, aux, True, ln)
    for T in (1, 2, 7, 11, 30):
        for seed in range(3):
            sim = Sim(seed + 31 * T, T, p["kblocks"], stages=p["stages"], kgroup=p["kgroup"], nacc=p["nacc"], nbuf=p["nbuf"],
                      aux=aux, lnf=ln)
            assert sim.two == bool(p["two_issuers"])
            sim.run()

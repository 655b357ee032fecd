"""Host-side pieces of bench.py that do not need a GPU: the committed ncu traffic lookup, the algorithmic byte counts
 and the clock sampler's behaviour on a box without NVML."""
import ctypes

import bench


def test_traffic_lookup_reads_the_committed_capture():
    t = bench.measured_traffic("srfrd_gemm_tn")
    assert isinstance(t, int) and 1e6 < t < 200e6           # DRAM bytes per launch of the dominant kernel at C2 (packed rows: ~7 MB)
    assert bench.measured_traffic("srfrd_catalogue_topk") > 1e9
    assert bench.measured_traffic("no_such_kernel") is None
    for k in ("srfrd_score_loss_fused@C3", "srfrd_embed_bwd@C3", "srfrd_adam_step@C3", "srfrd_embed_ln_fwd@C3"):
        assert bench.measured_traffic(k) > 1e9              # catalogue-scale captures of K4 / K5 / K7 / K1


def test_algorithmic_bytes_of_the_gemm_variants():
    from srfrd_b200._lib import GemmEpilogue
    M, N, K = 204800, 80, 80

    def args(**kw):
        ep = GemmEpilogue()
        for k, v in kw.items():
            setattr(ep, k, v)
        return [None, 0, None, 0, M, N, K, ctypes.byref(ep), None]
    base = M * K * 2 + N * K * 2 + M * N * 2
    b, f = bench.algorithmic_bytes("srfrd_gemm_tn", args(out_bf16=1), 0.1)
    assert b == base and f == 2.0 * M * N * K
    b, _ = bench.algorithmic_bytes("srfrd_gemm_tn", args(out_bf16=1, residual=1, row_ids=1), 0.1)
    assert b == base + M * N * 2 + M * 8
    b, _ = bench.algorithmic_bytes("srfrd_gemm_tn", args(out_bf16=1, residual=1, ln_out_bf16=1, ln_stats=1), 0.1)
    assert b == base + M * N * 2 + M * N * 2 + M * 8          # + LayerNorm output and (mean, rstd)
    b, _ = bench.algorithmic_bytes("srfrd_gemm_wgrad", [None, 0, None, 0, M, 80, 80], 0.1)
    assert b == M * 160 * 2 + 80 * 80 * 4


def test_clock_sampler_degrades_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"} and out["samples"] == 0

// Two chained GEMMs in one kernel: out = epilogue2( epilogue1(A W1^T) W2^T ), both weights square (N = K <= 128).
//
// replaces: the two 1x1 Conv1d of PointWiseFeedForward (SRFR_model.py:41,44,47-51) -- forward
//             h1 = relu(drop1(y W1^T + b1));  x' = (drop2(h1 W2^T + b2) + y) * (id != 0)   [+ the next LayerNorm]
//           and their data-gradient pair in backward
//             da1 = drop1'((dz2 W2) ) * (h1 > 0);  dy = da1 W1 + dz
// which the first version ran as two srfrd_gemm_tn launches each.  On the packed token layout a launch moves ~5 MB and
// costs ~10 us of fixed latency (launch, TMEM allocation, first TMA round trip, store, release): the step is a chain of
// such latencies, so the lever is the NUMBER of launches.  Here the intermediate tile never leaves the SM on its way to
// the second MMA: epilogue 1 writes it to shared memory as the K-major 128-byte-swizzled A operand of MMA 2 (and a TMA
// store copies the same tile to HBM, because backward / the weight gradients need it).
//
// One CTA per 128-row tile (grid-stride for larger inputs), two CTAs per SM: 104 KB of shared memory (both weights
// resident, A tile, intermediate tile) and 256 TMEM columns each, so at C2's 222 tiles every tile has its own CTA and the
// whole kernel is ONE pass of: TMA (weights + A) -> MMA 1 -> epilogue 1 -> MMA 2 -> epilogue 2 -> TMA stores.
// Warps 0..3: epilogue, thread = row = TMEM lane (row statistics of the fused LayerNorm need no exchange);
// warp 4: TMA producer + MMA issuer.  The stage-2 result is written in place over the A tile (which is the residual in the
// forward and, without dropout, in the backward), the LayerNorm output over the intermediate tile.
#include <stdlib.h>

#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

static constexpr int MLP_THREADS = 160;
static constexpr int MT = 128;                    // rows per tile
static constexpr int MBLK = MT * 64 * 2;          // [128 x 64] bf16 block, 16 KB

struct MlpParams {
  int M, N, kb;                                   // rows (capacity), output width, 64-column blocks of the output width
  int K, kbk;                                     // width of A (two-stage mode: K == N) and its 64-column blocks
  int single;                                     // one GEMM only: out = epilogue2(A W1^T)  (the lean srfrd_gemm_tn replacement)
  int relu2; const bf16* gate2; int ldg2;         // stage-2 ReLU / gate (single mode)
  const int* rows_dev;
  const float *bias1, *bias2;
  const bf16* gate; int ldg;                      // stage 1: v = gate > 0 ? v : 0
  int relu1;
  uint64_t drop_seed; const float* drop_step;
  uint32_t thresh1, thresh2, stream1, stream2; float scale1, scale2;
  const bf16* residual; int ldr;                  // stage 2: += residual; null = the A tile itself
  int residual_is_a;
  const int64_t* row_ids;
  const float *ln_w, *ln_b; float* ln_stats; float ln_eps; int has_ln;
};

__device__ __forceinline__ uint32_t swz_chunk(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ void mlp_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void mlp_tma_store(const CUtensorMap* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void unpack8_bf16(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

__global__ void __launch_bounds__(MLP_THREADS, 2)
mlp2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
            const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmMid,
            const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmLn, MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int wblk = (p.N * 128 + 1023) & ~1023;     // one 64-column K block of a weight: N rows x 128 B
  const int ablk = max(p.kb, p.kbk);                // the A tile also receives the result in place
  uint8_t* sW1 = smem;
  uint8_t* sW2 = sW1 + p.kbk * wblk;
  uint8_t* sA = sW2 + (p.single ? 0 : p.kb) * wblk;
  uint8_t* sM = sA + ablk * MBLK;
  float* sb1 = reinterpret_cast<float*>(sM + p.kb * MBLK);
  float* sb2 = sb1 + 128;
  float* slw = sb2 + 128;
  float* slb = slw + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(slb + 128);
  uint64_t *w_full = bars, *a_full = bars + 1, *a_free = bars + 2, *acc1_full = bars + 3, *mid_ready = bars + 4,
           *acc2_full = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmMid);
    tma_prefetch_desc(&tmOut);
    if (p.has_ln) tma_prefetch_desc(&tmLn);
    mbar_init(w_full, 1); mbar_init(a_full, 1); mbar_init(a_free, 1); mbar_init(acc1_full, 1); mbar_init(mid_ready, 1);
    mbar_init(acc2_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  pdl_prologue_done();
  const int M = p.rows_dev ? min(p.M, __ldg(p.rows_dev)) : p.M;
  const int n_tiles = (M + MT - 1) / MT;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    sb1[i] = (p.bias1 && i < p.N) ? __ldg(p.bias1 + i) : 0.f;
    sb2[i] = (p.bias2 && i < p.N) ? __ldg(p.bias2 + i) : 0.f;
    slw[i] = (p.has_ln && i < p.N) ? __ldg(p.ln_w + i) : 0.f;
    slb[i] = (p.has_ln && i < p.N) ? __ldg(p.ln_b + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tAcc1 = tmem_base, tAcc2 = tmem_base + 128;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer + MMA issuer
    if ((int)blockIdx.x < n_tiles && elect_one()) {
      mbar_expect_tx(w_full, (p.kbk + (p.single ? 0 : p.kb)) * p.N * 128);
      for (int kb = 0; kb < p.kbk; ++kb) tma_load_2d(sW1 + kb * wblk, &tmW1, w_full, kb * 64, 0, SRFRD_EVICT_LAST);
      if (!p.single)
        for (int kb = 0; kb < p.kb; ++kb) tma_load_2d(sW2 + kb * wblk, &tmW2, w_full, kb * 64, 0, SRFRD_EVICT_LAST);
    }
    __syncwarp();
    const uint32_t idesc = umma_idesc_bf16(MT, p.N, 0, 0);
    const uint64_t dA = umma_smem_desc(smem_u32(sA), 0, 1024), dM = umma_smem_desc(smem_u32(sM), 0, 1024);
    const uint64_t dW1 = umma_smem_desc(smem_u32(sW1), 0, 1024), dW2 = umma_smem_desc(smem_u32(sW2), 0, 1024);
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ph ^= 1) {
      mbar_wait(a_free, ph ^ 1);                       // the previous tile's results have left the A / intermediate tiles
      if (elect_one()) {
        mbar_expect_tx(a_full, p.kbk * MBLK);
        for (int kb = 0; kb < p.kbk; ++kb) tma_load_2d(sA + kb * MBLK, &tmA, a_full, kb * 64, tile * MT, SRFRD_EVICT_FIRST);
      }
      __syncwarp();
      if (tile == (int)blockIdx.x) mbar_wait(w_full, 0);
      mbar_wait(a_full, ph);
      tc_fence_after();
      if (elect_one()) {
        for (int kb = 0; kb < p.kbk; ++kb) {
          const int ksteps = min(4, (p.K - kb * 64 + 15) / 16);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(p.single ? tAcc2 : tAcc1, dA + (uint64_t)(kb * (MBLK >> 4) + 2 * k),
                      dW1 + (uint64_t)(kb * (wblk >> 4) + 2 * k), idesc, (kb | k) != 0);
        }
        umma_commit(p.single ? acc2_full : acc1_full);
      }
      __syncwarp();
      if (p.single) continue;
      mbar_wait(mid_ready, ph);                        // epilogue 1 has written the intermediate tile (A operand of MMA 2)
      tc_fence_after();
      if (elect_one()) {
        for (int kb = 0; kb < p.kb; ++kb) {
          const int ksteps = min(4, (p.N - kb * 64 + 15) / 16);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tAcc2, dM + (uint64_t)(kb * (MBLK >> 4) + 2 * k), dW2 + (uint64_t)(kb * (wblk >> 4) + 2 * k), idesc, (kb | k) != 0);
        }
        umma_commit(acc2_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogues: thread = row
    const int r = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint64_t seed = (p.thresh1 | p.thresh2) ? mix_seed(p.drop_seed, p.drop_step) : 0;
    const bool issuer = threadIdx.x == 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ph ^= 1) {
      const int64_t row = (int64_t)tile * MT + r;
      const bool rok = row < M;                         // (dense layout: the last tile may be partial; TMA clips its stores)
      // ---- stage 1
      if (!p.single) {
      mbar_wait(acc1_full, ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < p.N; c += 16) {
        uint32_t raw[16];
        tmem_ld16(tAcc1 + lane_off + c, raw);
        uint4 g0 = make_uint4(0, 0, 0, 0), g1 = g0;
        if (p.gate && rok) {
          const uint4* gp = reinterpret_cast<const uint4*>(p.gate + row * p.ldg + c);
          g0 = __ldg(gp); g1 = __ldg(gp + 1);
        }
        tmem_ld_wait();
        float v[16], gv[16];
        if (p.gate) { unpack8_bf16(g0, gv); unpack8_bf16(g1, gv + 8); }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = __uint_as_float(raw[i]) + sb1[c + i];
          if (p.thresh1)
            x = dropout_keep(seed, p.stream1, (uint64_t)row * (uint64_t)p.N + (uint64_t)(c + i), p.thresh1) ? x * p.scale1 : 0.f;
          if (p.relu1) x = fmaxf(x, 0.f);
          if (p.gate) x = gv[i] > 0.f ? x : 0.f;
          v[i] = x;
        }
        uint8_t* blk = sM + (c >> 6) * MBLK;
        const int j = (c & 63) >> 3;
        *reinterpret_cast<uint4*>(blk + swz_chunk(r, j)) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        *reinterpret_cast<uint4*>(blk + swz_chunk(r, j + 1)) =
            make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
      }
      tc_fence_before();
      fence_proxy_async();                            // generic-proxy writes -> visible to the MMA and the TMA store
      mlp_bar_sync();
      if (issuer) {
        mbar_arrive(mid_ready);
        for (int kb = 0; kb < p.kb; ++kb) mlp_tma_store(&tmMid, sM + kb * MBLK, kb * 64, tile * MT);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      }
      // ---- stage 2
      const float rowm = (p.row_ids && rok) ? ((__ldg(p.row_ids + row) != 0) ? 1.f : 0.f) : 1.f;
      mbar_wait(acc2_full, ph);
      tc_fence_after();
      float ln_sum = 0.f, ln_sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < p.N; c += 16) {
        uint32_t raw[16];
        tmem_ld16(tAcc2 + lane_off + c, raw);
        uint8_t* blk = sA + (c >> 6) * MBLK;
        const int j = (c & 63) >> 3;
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
        if (p.residual_is_a && c < p.K) {
          r0 = *reinterpret_cast<const uint4*>(blk + swz_chunk(r, j));
          r1 = *reinterpret_cast<const uint4*>(blk + swz_chunk(r, j + 1));
        } else if (p.residual && rok) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row * p.ldr + c);
          r0 = __ldg(rp); r1 = __ldg(rp + 1);
        }
        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
        if (p.gate2 && rok) {
          const uint4* gp = reinterpret_cast<const uint4*>(p.gate2 + row * p.ldg2 + c);
          q0 = __ldg(gp); q1 = __ldg(gp + 1);
        }
        tmem_ld_wait();
        float v[16], rv[16], gv2[16];
        unpack8_bf16(r0, rv); unpack8_bf16(r1, rv + 8);
        if (p.gate2) { unpack8_bf16(q0, gv2); unpack8_bf16(q1, gv2 + 8); }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = __uint_as_float(raw[i]) + sb2[c + i];
          if (p.thresh2)
            x = dropout_keep(seed, p.stream2, (uint64_t)row * (uint64_t)p.N + (uint64_t)(c + i), p.thresh2) ? x * p.scale2 : 0.f;
          if (p.relu2) x = fmaxf(x, 0.f);
          if (p.gate2) x = gv2[i] > 0.f ? x : 0.f;
          x += rv[i];
          v[i] = x * rowm;
        }
        const uint4 p0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        const uint4 p1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        *reinterpret_cast<uint4*>(blk + swz_chunk(r, j)) = p0;
        *reinterpret_cast<uint4*>(blk + swz_chunk(r, j + 1)) = p1;
        if (p.has_ln) {                                // statistics of the ROUNDED values (what a LayerNorm kernel would read)
          float q[16];
          unpack8_bf16(p0, q); unpack8_bf16(p1, q + 8);
#pragma unroll
          for (int i = 0; i < 16; ++i) { ln_sum += q[i]; ln_sq = fmaf(q[i], q[i], ln_sq); }
        }
      }
      tc_fence_before();
      if (p.has_ln) {
        // the intermediate tile becomes the LayerNorm staging buffer: its TMA store must have finished READING it
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mlp_bar_sync();
        const float invn = 1.f / (float)p.N;
        const float mean = ln_sum * invn;
        const float var = fmaxf(ln_sq * invn - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.ln_eps), nb = -mean * rstd;
#pragma unroll 1
        for (int c = 0; c < p.N; c += 8) {
          const int j = (c & 63) >> 3;
          const uint4 xv = *reinterpret_cast<const uint4*>(sA + (c >> 6) * MBLK + swz_chunk(r, j));
          float q[8], y[8];
          unpack8_bf16(xv, q);
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = fmaf(fmaf(q[i], rstd, nb), slw[c + i], slb[c + i]);
          *reinterpret_cast<uint4*>(sM + (c >> 6) * MBLK + swz_chunk(r, j)) =
              make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
        }
        if (p.ln_stats && rok) *reinterpret_cast<float2*>(p.ln_stats + 2 * row) = make_float2(mean, rstd);
      }
      fence_proxy_async();
      mlp_bar_sync();
      if (issuer) {
        for (int kb = 0; kb < p.kb; ++kb) mlp_tma_store(&tmOut, sA + kb * MBLK, kb * 64, tile * MT);
        if (p.has_ln)
          for (int kb = 0; kb < p.kb; ++kb) mlp_tma_store(&tmLn, sM + kb * MBLK, kb * 64, tile * MT);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(a_free);
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // stores complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

}  // namespace srfrd

using namespace srfrd;

static int mlp_launch(MlpParams& p, const void* A, int lda, const void* W1, int ldw1, const void* W2, int ldw2, void* mid_out,
                      int ldm, void* out, int ldc, void* ln_out, int ld_ln, cudaStream_t stream) {
  p.kb = (p.N + 63) / 64; p.kbk = (p.K + 63) / 64; p.rows_dev = row_limit();
  p.has_ln = ln_out != nullptr;
  CUtensorMap tmA, tmW1, tmW2, tmMid, tmOut, tmLn;
  if (int rc = make_tmap_bf16_2d(&tmA, A, p.M, p.K, lda, MT, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmW1, W1, p.N, p.K, ldw1, p.N, 64)) return rc;
  tmW2 = tmW1; tmMid = tmA;
  if (!p.single) {
    if (int rc = make_tmap_bf16_2d(&tmW2, W2, p.N, p.N, ldw2, p.N, 64)) return rc;
    if (int rc = make_tmap_bf16_2d(&tmMid, mid_out, p.M, p.N, ldm, MT, 64)) return rc;
  }
  if (int rc = make_tmap_bf16_2d(&tmOut, out, p.M, p.N, ldc, MT, 64)) return rc;
  tmLn = tmOut;
  if (p.has_ln) if (int rc = make_tmap_bf16_2d(&tmLn, ln_out, p.M, p.N, ld_ln, MT, 64)) return rc;
  const int wblk = (p.N * 128 + 1023) & ~1023;
  const int ablk = p.kb > p.kbk ? p.kb : p.kbk;
  const size_t smem = 1024 + (size_t)(p.kbk + (p.single ? 0 : p.kb)) * wblk + (size_t)(ablk + p.kb) * MBLK + 4 * 128 * sizeof(float) + 128;
  static bool attr = false;
  if (!attr) {
    SRFRD_CUDA(cudaFuncSetAttribute(mlp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  SRFRD_REQUIRE(smem <= 200 * 1024, "mlp2_tn: shared memory %zu too large", smem);
  int64_t grid = ((int64_t)p.M + MT - 1) / MT;
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  SRFRD_CUDA(launch_pdl(mlp2_kernel, dim3((unsigned)grid), dim3(MLP_THREADS), smem, stream, tmA, tmW1, tmW2, tmMid, tmOut, tmLn, p));
  SRFRD_LAUNCH_CHECK();
  return 0;
}

namespace srfrd {
// The lean single-GEMM path of srfrd_gemm_tn (see gemm.cu): one CTA per 128-row tile, two CTAs per SM, no pipeline -- for
// inputs of a few hundred tiles (the packed token layout) where the persistent kernel's fixed cost dominates.
bool gemm_small_eligible(int M, int N, int K, const srfrd_gemm_epilogue_t* ep, int lda, int ldb) {
  if (!ep->out_bf16 || ep->out_f32 || N > 128 || N % 16 || K > 128 || K % 8 || lda % 8 || ldb % 8 || ep->ldc % 8) return false;
  if (((uintptr_t)ep->out_bf16 & 15) || (ep->residual && (ep->ldr % 8 || ((uintptr_t)ep->residual & 15)))) return false;
  if (ep->gate && (ep->ldg % 8 || ((uintptr_t)ep->gate & 15))) return false;
  if (ep->residual && ep->gate) return false;
  if (ep->ln_out_bf16 && (!ep->ln_w || !ep->ln_b || ep->ld_ln % 8 || ((uintptr_t)ep->ln_out_bf16 & 15))) return false;
  // Measured and NOT adopted (opt-in, SRFRD_LEAN_GEMM=1): at C2's 222 packed tiles the lean kernel takes 14.8 us per launch
  // against 13.7 us for the persistent pipelined kernel, and at C5's 590 tiles it is slower (two un-pipelined tiles per
  // CTA: step 0.92 vs 0.85 ms).  The fixed cost of a launch is NOT in the pipeline's structure (it is launch + TMEM
  // allocation + first TMA round trip + store drain, common to both); only fewer launches help (srfrd_mlp2_tn).
  { const char* e = getenv("SRFRD_LEAN_GEMM"); if (!e || e[0] != '1') return false; }
  return row_limit() != nullptr || M <= 32768;
}
int gemm_small_launch(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const srfrd_gemm_epilogue_t* ep,
                      cudaStream_t stream) {
  MlpParams p = {};
  p.M = M; p.N = N; p.K = K; p.single = 1;
  p.bias2 = ep->bias; p.relu2 = ep->relu; p.gate2 = (const bf16*)ep->gate; p.ldg2 = ep->ldg;
  p.residual = (const bf16*)ep->residual; p.ldr = ep->ldr;
  p.residual_is_a = (ep->residual == A && ep->ldr == lda && N <= K) ? 1 : 0;
  if (p.residual_is_a) p.residual = nullptr;
  p.row_ids = ep->row_ids;
  p.drop_seed = ep->drop_seed; p.drop_step = ep->drop_step; p.stream2 = ep->drop_stream;
  p.thresh2 = ep->drop_p > 0.f ? (uint32_t)((double)ep->drop_p * 4294967296.0) : 0;
  p.scale2 = ep->drop_p > 0.f ? 1.f / (1.f - ep->drop_p) : 1.f;
  p.scale1 = 1.f;
  p.ln_w = ep->ln_w; p.ln_b = ep->ln_b; p.ln_stats = ep->ln_stats; p.ln_eps = ep->ln_eps;
  return mlp_launch(p, A, lda, B, ldb, nullptr, 0, nullptr, 0, ep->out_bf16, ep->ldc, ep->ln_out_bf16, ep->ld_ln, stream);
}
}  // namespace srfrd

extern "C" int srfrd_mlp2_tn(const void* A, int lda, const void* W1, int ldw1, const void* W2, int ldw2, int M, int N,
                             const srfrd_mlp2_t* ep, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRFRD_REQUIRE(A && W1 && W2 && ep && ep->mid_out && ep->out, "mlp2_tn: null pointer");
  SRFRD_REQUIRE(M > 0 && N >= 16 && N <= 128 && N % 16 == 0, "mlp2_tn: N = K must be a multiple of 16 up to 128 (got %d)", N);
  SRFRD_REQUIRE(lda % 8 == 0 && ldw1 % 8 == 0 && ldw2 % 8 == 0 && ep->ldm % 8 == 0 && ep->ldc % 8 == 0 && lda >= N && ldw1 >= N &&
                    ldw2 >= N && ep->ldm >= N && ep->ldc >= N, "mlp2_tn: bad leading dimensions");
  SRFRD_REQUIRE(!ep->gate || (ep->ldg % 8 == 0 && ((uintptr_t)ep->gate & 15) == 0), "mlp2_tn: bad gate");
  SRFRD_REQUIRE(!ep->residual || (ep->ldr % 8 == 0 && ((uintptr_t)ep->residual & 15) == 0), "mlp2_tn: bad residual");
  SRFRD_REQUIRE(!ep->ln_out || (ep->ln_w && ep->ln_b && ep->ld_ln % 8 == 0 && ep->ld_ln >= N), "mlp2_tn: bad LayerNorm arguments");
  SRFRD_REQUIRE(ep->drop1_p >= 0.f && ep->drop1_p < 1.f && ep->drop2_p >= 0.f && ep->drop2_p < 1.f, "mlp2_tn: dropout p must be in [0, 1)");
  MlpParams p = {};
  p.M = M; p.N = N; p.K = N; p.single = 0;
  p.bias1 = ep->bias1; p.bias2 = ep->bias2; p.gate = (const bf16*)ep->gate; p.ldg = ep->ldg; p.relu1 = ep->relu1;
  p.drop_seed = ep->drop_seed; p.drop_step = ep->drop_step; p.stream1 = ep->drop1_stream; p.stream2 = ep->drop2_stream;
  p.thresh1 = ep->drop1_p > 0.f ? (uint32_t)((double)ep->drop1_p * 4294967296.0) : 0;
  p.thresh2 = ep->drop2_p > 0.f ? (uint32_t)((double)ep->drop2_p * 4294967296.0) : 0;
  p.scale1 = ep->drop1_p > 0.f ? 1.f / (1.f - ep->drop1_p) : 1.f;
  p.scale2 = ep->drop2_p > 0.f ? 1.f / (1.f - ep->drop2_p) : 1.f;
  p.residual = (const bf16*)ep->residual; p.ldr = ep->ldr;
  p.residual_is_a = (ep->residual == nullptr && ep->residual_is_a) || (ep->residual == A && ep->ldr == lda);
  if (p.residual_is_a) p.residual = nullptr;
  p.row_ids = ep->row_ids;
  p.ln_w = ep->ln_w; p.ln_b = ep->ln_b; p.ln_stats = ep->ln_stats; p.ln_eps = ep->ln_eps;
  return mlp_launch(p, A, lda, W1, ldw1, W2, ldw2, ep->mid_out, ep->ldm, ep->out, ep->ldc, ep->ln_out, ep->ld_ln, stream);
}

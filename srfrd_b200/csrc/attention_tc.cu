// Causal self-attention on tcgen05 for maxlen <= 128 (SRFR_model.py:112).
//
// A tile is a window of 128 consecutive tokens that holds spt = floor(128 / L) whole sequences, so the
// queries AND the keys/values of those sequences live in the same window:
//     S[128 x 128] = Q K^T   (tcgen05, fp32 in TMEM)   -> per-row causal/same-sequence softmax in registers
//     O[128 x hd ] = P V     (P written to smem as a bf16 K-major operand, V read MN-major from its TMA tile)
// One thread owns one query row (TMEM lane), so max / sum / delta need no cross-thread reduction, and the
// (L x L) probabilities never touch HBM.  Backward recomputes S and P and runs five MMAs per tile:
//     S = Q K^T, dP = dO V^T, dQ = dS K, dK = dS^T Q, dV = P^T dO
// where the SAME 128-byte-swizzled [128 tokens x 64 features] TMA tile serves as a K-major operand (rows = M/N)
// and as an MN-major operand (rows = K) -- only the descriptor changes.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 softmax + epilogue.
#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

int attn_fwd_simt(const void* q, int ldq, const void* k, const void* v, int ldkv, void* o, int ldo, int64_t B, int L,
                  int H, int heads, float drop_p, uint64_t seed, uint32_t stream_id, const float* drop_step, void* stream);
int attn_bwd_simt(const void* dout, int lddo, const void* q, int ldq, const void* k, const void* v, int ldkv, void* dq,
                  int lddq, void* dk, void* dv, int lddkv, int64_t B, int L, int H, int heads, float drop_p,
                  uint64_t seed, uint32_t stream_id, const float* drop_step, void* stream);

static constexpr int TILE = 128;
static constexpr int OPB = TILE * 64 * 2;     // operand block: 128 rows x 64 bf16 = 16 KB
static constexpr int ATC_THREADS = 192;

struct AttnTc {
  int64_t T;
  int L, hd, heads, spt, n_tiles, kblocks;
  float scale, scale_log2e;
  bf16* o; int ldo;                       // forward output
  bf16 *dq, *dk, *dv; int lddq, lddkv;    // backward outputs
  uint64_t drop_seed; uint32_t drop_thresh, drop_stream; float drop_scale; const float* drop_step;
};

// byte offset of element (row r, column c) inside a [128 x 64] bf16 block with the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128(int r, int c) {
  return (uint32_t)(r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1)));
}

struct RowInfo {
  int r; int64_t t; bool owned; int jlo, jhi, l; int64_t bh;
};
__device__ __forceinline__ RowInfo row_info(const AttnTc& p, int tile, int h, int quarter, int lane) {
  RowInfo ri;
  ri.r = quarter * 32 + lane;
  ri.t = (int64_t)tile * p.spt * p.L + ri.r;
  ri.owned = ri.r < p.spt * p.L && ri.t < p.T;
  const int sidx = ri.r / p.L;
  ri.l = ri.r - sidx * p.L;
  ri.jlo = sidx * p.L;
  ri.jhi = ri.jlo + ri.l;
  ri.bh = (ri.t / p.L) * p.heads + h;
  return ri;
}
__device__ __forceinline__ float drop_f(const AttnTc& p, const RowInfo& ri, int col) {
  if (!p.drop_thresh) return 1.f;
  const uint64_t idx = ((uint64_t)ri.bh * p.L + ri.l) * p.L + (col - ri.jlo);
  return dropout_keep(p.drop_seed, p.drop_stream, idx, p.drop_thresh) ? p.drop_scale : 0.f;
}

// row max and sum of exp over the causal window of this thread's row (two passes over the S accumulator)
__device__ __forceinline__ void row_softmax_stats(uint32_t tS, const AttnTc& p, const RowInfo& ri, float& m, float& sum) {
  m = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t raw[32];
    tmem_ld32(tS + c * 32, raw);
    tmem_ld_wait();
    const int lo = ri.jlo - c * 32;
    const unsigned span = (unsigned)(ri.jhi - ri.jlo);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if ((unsigned)(j - lo) <= span) m = fmaxf(m, __uint_as_float(raw[j]));
  }
  sum = 0.f;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t raw[32];
    tmem_ld32(tS + c * 32, raw);
    tmem_ld_wait();
    const int lo = ri.jlo - c * 32;
    const unsigned span = (unsigned)(ri.jhi - ri.jlo);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if ((unsigned)(j - lo) <= span) sum += exp2f((__uint_as_float(raw[j]) - m) * p.scale_log2e);
  }
}

__device__ __forceinline__ void store_row32(uint8_t* blk_base, int r, int c, const float* v) {
  // 32 consecutive keys [c*32, c*32+32) of row r -> 4 x 16 B into the swizzled K-major block pair
  uint8_t* blk = blk_base + (c >> 1) * OPB;
  const int col0 = (c & 1) * 32;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint4 w = make_uint4(pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                               pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
    *reinterpret_cast<uint4*>(blk + sw128(r, col0 + 8 * u)) = w;
  }
}

__device__ __forceinline__ void store_out_row(bf16* dst, uint32_t tacc, int hd, float mul) {
  for (int c = 0; c < hd; c += 16) {
    uint32_t raw[16];
    tmem_ld16(tacc + c, raw);
    tmem_ld_wait();
    if (dst) {
      uint4* o = reinterpret_cast<uint4*>(dst + c);
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) * mul;
      o[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      o[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                        pack_bf16x2(v[14], v[15]));
    }
  }
}

// =============================================================================================== forward
__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, AttnTc p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int opbytes = p.kblocks * OPB;
  uint8_t* Qs = smem;
  uint8_t* Ks = Qs + opbytes;
  uint8_t* Vs = Ks + opbytes;
  uint8_t* Ps = Vs + opbytes;                     // 2 key blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + 2 * OPB);
  uint64_t *qk_full = bars, *qk_empty = bars + 1, *v_full = bars + 2, *v_empty = bars + 3, *s_full = bars + 4,
           *p_full = bars + 5, *o_full = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_tiles * p.heads;
  if (p.drop_thresh) p.drop_seed = mix_seed(p.drop_seed, p.drop_step);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(qk_full, 1); mbar_init(qk_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  // warps 0 / 1 stay converged; issuing instructions sit under elect.sync (uniform-register operands)
  if (warp == 0) {
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ph ^= 1) {
      const int tile = it / p.heads, h = it % p.heads;
      const int row0 = tile * p.spt * p.L, col0 = h * p.hd;
      mbar_wait(qk_empty, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(qk_full, 2 * opbytes);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tma_load_2d(Qs + kb * OPB, &tmQ, qk_full, col0 + kb * 64, row0, SRFRD_EVICT_FIRST);
          tma_load_2d(Ks + kb * OPB, &tmK, qk_full, col0 + kb * 64, row0, SRFRD_EVICT_FIRST);
        }
      }
      __syncwarp();
      mbar_wait(v_empty, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(v_full, opbytes);
        for (int kb = 0; kb < p.kblocks; ++kb)
          tma_load_2d(Vs + kb * OPB, &tmV, v_full, col0 + kb * 64, row0, SRFRD_EVICT_FIRST);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idS = umma_idesc_bf16(TILE, TILE, 0, 0);
    const uint32_t idO = umma_idesc_bf16(TILE, p.hd, 0, 1);
    const uint64_t dQ = umma_smem_desc(smem_u32(Qs), 0, 1024), dK = umma_smem_desc(smem_u32(Ks), 0, 1024);
    const uint64_t dP = umma_smem_desc(smem_u32(Ps), 0, 1024), dV = umma_smem_desc(smem_u32(Vs), OPB, 1024);
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ph ^= 1) {
      mbar_wait(qk_full, ph);
      tc_fence_after();
      if (elect_one()) {
        for (int kb = 0; kb < p.kblocks; ++kb) {
          const int ksteps = min(4, (p.hd - kb * 64) / 16);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tS, dQ + (uint64_t)(kb * (OPB >> 4) + 2 * k), dK + (uint64_t)(kb * (OPB >> 4) + 2 * k), idS, (kb | k) != 0);
        }
        umma_commit(qk_empty);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, ph);
      mbar_wait(v_full, ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)     // K = 128 keys; V tile read MN-major (rows = K): 16 rows = 2048 B
          umma_bf16(tO, dP + (uint64_t)((ks >> 2) * (OPB >> 4) + (ks & 3) * 2), dV + (uint64_t)(ks * 128), idO, ks != 0);
        umma_commit(v_empty);
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ph ^= 1) {
      const int tile = it / p.heads, h = it % p.heads;
      RowInfo ri = row_info(p, tile, h, quarter, lane);
      if (!ri.owned) { ri.jlo = 1 << 20; ri.jhi = ri.jlo; }         // empty window; TMEM loads stay warp-collective
      mbar_wait(s_full, ph);
      tc_fence_after();
      float m, sum;
      row_softmax_stats(tS + lane_off, p, ri, m, sum);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld32(tS + lane_off + c * 32, raw);
        tmem_ld_wait();
        float pv[32];
        const int lo = ri.jlo - c * 32;
        const unsigned span = (unsigned)(ri.jhi - ri.jlo);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float e = 0.f;
          if ((unsigned)(j - lo) <= span) {
            e = exp2f((__uint_as_float(raw[j]) - m) * p.scale_log2e);      // un-normalised; O is scaled by 1/sum
            if (p.drop_thresh) e *= drop_f(p, ri, c * 32 + j);
          }
          pv[j] = e;
        }
        store_row32(Ps, ri.r, c, pv);
      }
      tc_fence_before();
      fence_proxy_async();                 // generic-proxy smem writes -> visible to the MMA (async proxy)
      mbar_arrive(p_full);
      mbar_wait(o_full, ph);
      tc_fence_after();
      bf16* dst = ri.owned ? p.o + ri.t * p.ldo + h * p.hd : nullptr;
      store_out_row(dst, tO + lane_off, p.hd, 1.f / sum);
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// =============================================================================================== backward
__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, AttnTc p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int opbytes = p.kblocks * OPB;
  uint8_t* Qs = smem;
  uint8_t* Ks = Qs + opbytes;
  uint8_t* Vs = Ks + opbytes;
  uint8_t* Ds = Vs + opbytes;                    // dO
  uint8_t* Ps = Ds + opbytes;                    // P * dropout mask, 2 key blocks
  uint8_t* Gs = Ps + 2 * OPB;                    // dS, 2 key blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(Gs + 2 * OPB);
  uint64_t *in_full = bars, *in_empty = bars + 1, *sdp_full = bars + 2, *ds_full = bars + 3, *out_full = bars + 4,
           *out_empty = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_tiles * p.heads;
  if (p.drop_thresh) p.drop_seed = mix_seed(p.drop_seed, p.drop_step);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(in_full, 1); mbar_init(in_empty, 1); mbar_init(sdp_full, 1); mbar_init(ds_full, 128); mbar_init(out_full, 1);
    mbar_init(out_empty, 128);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // S and dP live in columns [0,128) and [128,256); once the softmax backward has consumed them the same
  // columns receive dK and dV, and dQ goes to [256, 256+hd).
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDK = tmem_base, tDV = tmem_base + 128, tDQ = tmem_base + 256;

  if (warp == 0) {
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ph ^= 1) {
      const int tile = it / p.heads, h = it % p.heads;
      const int row0 = tile * p.spt * p.L, col0 = h * p.hd;
      mbar_wait(in_empty, ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(in_full, 4 * opbytes);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          tma_load_2d(Qs + kb * OPB, &tmQ, in_full, col0 + kb * 64, row0, SRFRD_EVICT_FIRST);
          tma_load_2d(Ks + kb * OPB, &tmK, in_full, col0 + kb * 64, row0, SRFRD_EVICT_FIRST);
          tma_load_2d(Vs + kb * OPB, &tmV, in_full, col0 + kb * 64, row0, SRFRD_EVICT_FIRST);
          tma_load_2d(Ds + kb * OPB, &tmDO, in_full, col0 + kb * 64, row0, SRFRD_EVICT_FIRST);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idKK = umma_idesc_bf16(TILE, TILE, 0, 0);      // S, dP : both operands K-major
    const uint32_t idKM = umma_idesc_bf16(TILE, p.hd, 0, 1);      // dQ    : A = dS K-major, B = K MN-major
    const uint32_t idMM = umma_idesc_bf16(TILE, p.hd, 1, 1);      // dK, dV: A = dS^T / P^T MN-major, B MN-major
    // K-major views (+2 per 16-column K step, +OPB>>4 per 64-column block) and MN-major views (+128 per 16-row K step)
    const uint64_t kQ = umma_smem_desc(smem_u32(Qs), 0, 1024), kK = umma_smem_desc(smem_u32(Ks), 0, 1024);
    const uint64_t kV = umma_smem_desc(smem_u32(Vs), 0, 1024), kD = umma_smem_desc(smem_u32(Ds), 0, 1024);
    const uint64_t kG = umma_smem_desc(smem_u32(Gs), 0, 1024);
    const uint64_t mK = umma_smem_desc(smem_u32(Ks), OPB, 1024), mQ = umma_smem_desc(smem_u32(Qs), OPB, 1024);
    const uint64_t mD = umma_smem_desc(smem_u32(Ds), OPB, 1024), mG = umma_smem_desc(smem_u32(Gs), OPB, 1024);
    const uint64_t mP = umma_smem_desc(smem_u32(Ps), OPB, 1024);
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ph ^= 1) {
      mbar_wait(in_full, ph);
      mbar_wait(out_empty, ph ^ 1);          // the previous item's dK/dV (same TMEM columns as S/dP) were read out
      tc_fence_after();
      if (elect_one()) {
        for (int kb = 0; kb < p.kblocks; ++kb) {
          const int ksteps = min(4, (p.hd - kb * 64) / 16);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t o = (uint64_t)(kb * (OPB >> 4) + 2 * k);
            umma_bf16(tS, kQ + o, kK + o, idKK, (kb | k) != 0);
            umma_bf16(tDP, kD + o, kV + o, idKK, (kb | k) != 0);
          }
        }
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(ds_full, ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ko = (uint64_t)((ks >> 2) * (OPB >> 4) + (ks & 3) * 2);   // K-major A: key block, 32 B per step
          const uint64_t mo = (uint64_t)(ks * 128);                                // MN-major: 16 rows = 2048 B
          umma_bf16(tDQ, kG + ko, mK + mo, idKM, ks != 0);      // dQ[i, :] += dS[i, keys] K[keys, :]
          umma_bf16(tDK, mG + mo, mQ + mo, idMM, ks != 0);      // dK[j, :] += dS[rows, j]^T Q[rows, :]
          umma_bf16(tDV, mP + mo, mD + mo, idMM, ks != 0);      // dV[j, :] += P[rows, j]^T dO[rows, :]
        }
        umma_commit(in_empty);
        umma_commit(out_full);
      }
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ph ^= 1) {
      const int tile = it / p.heads, h = it % p.heads;
      RowInfo ri = row_info(p, tile, h, quarter, lane);
      if (!ri.owned) { ri.jlo = 1 << 20; ri.jhi = ri.jlo; }         // empty window, loads stay warp-collective
      mbar_wait(sdp_full, ph);
      tc_fence_after();
      float m, sum;
      row_softmax_stats(tS + lane_off, p, ri, m, sum);
      const float inv = ri.owned ? 1.f / sum : 0.f;
      float delta = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {        // delta_i = sum_j P_ij * dP_ij * M_ij
        uint32_t rs[32], rg[32];
        tmem_ld32(tS + lane_off + c * 32, rs);
        tmem_ld32(tDP + lane_off + c * 32, rg);
        tmem_ld_wait();
        const int lo = ri.jlo - c * 32;
        const unsigned span = (unsigned)(ri.jhi - ri.jlo);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((unsigned)(j - lo) <= span) {
            const float pr = exp2f((__uint_as_float(rs[j]) - m) * p.scale_log2e) * inv;
            delta += pr * __uint_as_float(rg[j]) * drop_f(p, ri, c * 32 + j);
          }
      }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t rs[32], rg[32];
        tmem_ld32(tS + lane_off + c * 32, rs);
        tmem_ld32(tDP + lane_off + c * 32, rg);
        tmem_ld_wait();
        float pd[32], ds[32];
        const int lo = ri.jlo - c * 32;
        const unsigned span = (unsigned)(ri.jhi - ri.jlo);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a = 0.f, g = 0.f;
          if ((unsigned)(j - lo) <= span) {
            const float pr = exp2f((__uint_as_float(rs[j]) - m) * p.scale_log2e) * inv;
            const float mk = drop_f(p, ri, c * 32 + j);
            a = pr * mk;
            g = pr * (__uint_as_float(rg[j]) * mk - delta) * p.scale;      // S = scale * q k^T
          }
          pd[j] = a; ds[j] = g;
        }
        store_row32(Ps, ri.r, c, pd);
        store_row32(Gs, ri.r, c, ds);
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(ds_full);
      mbar_wait(out_full, ph);
      tc_fence_after();
      const bool own = ri.jlo != (1 << 20);
      store_out_row(own ? p.dq + ri.t * p.lddq + h * p.hd : nullptr, tDQ + lane_off, p.hd, 1.f);
      store_out_row(own ? p.dk + ri.t * p.lddkv + h * p.hd : nullptr, tDK + lane_off, p.hd, 1.f);
      store_out_row(own ? p.dv + ri.t * p.lddkv + h * p.hd : nullptr, tDV + lane_off, p.hd, 1.f);
      tc_fence_before();
      mbar_arrive(out_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static bool tc_supported(int L, int H, int heads, int ldq, int ldkv) {
  if (heads <= 0 || H % heads) return false;
  const int hd = H / heads;
  return L >= 1 && L <= TILE && hd % 16 == 0 && hd <= 128 && ldq % 8 == 0 && ldkv % 8 == 0;
}

static int fill(AttnTc& p, int64_t B, int L, int H, int heads, float drop_p, uint64_t seed, uint32_t stream_id,
                const float* drop_step) {
  p.T = B * L; p.L = L; p.heads = heads; p.hd = H / heads;
  p.spt = TILE / L;
  p.n_tiles = (int)((B + p.spt - 1) / p.spt);
  p.kblocks = (p.hd + 63) / 64;
  p.scale = 1.0f / sqrtf((float)p.hd);
  p.scale_log2e = p.scale * 1.4426950408889634f;
  p.drop_seed = seed; p.drop_stream = stream_id; p.drop_step = drop_step; p.drop_thresh = 0; p.drop_scale = 1.f;
  if (drop_p > 0.f) {
    SRFRD_REQUIRE(drop_p < 1.f, "attention: dropout p must be < 1");
    p.drop_thresh = (uint32_t)((double)drop_p * 4294967296.0);
    p.drop_scale = 1.f / (1.f - drop_p);
  }
  return 0;
}

}  // namespace srfrd

using namespace srfrd;

extern "C" int srfrd_attention_fwd(const void* q, int ldq, const void* k, const void* v, int ldkv, void* o, int ldo,
                                   int64_t B, int L, int H, int heads, float drop_p, uint64_t seed,
                                   uint32_t stream_id, const float* drop_step, void* stream) {
  SRFRD_REQUIRE(q && k && v && o, "attention_fwd: null pointer");
  if (B == 0 || L == 0) return 0;
  if (!tc_supported(L, H, heads, ldq, ldkv) || ldo % 8)
    return attn_fwd_simt(q, ldq, k, v, ldkv, o, ldo, B, L, H, heads, drop_p, seed, stream_id, drop_step, stream);
  AttnTc p = {};
  if (int rc = fill(p, B, L, H, heads, drop_p, seed, stream_id, drop_step)) return rc;
  p.o = (bf16*)o; p.ldo = ldo;
  CUtensorMap tmQ, tmK, tmV;
  if (int rc = make_tmap_bf16_2d(&tmQ, q, p.T, H, ldq, TILE, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmK, k, p.T, H, ldkv, TILE, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmV, v, p.T, H, ldkv, TILE, 64)) return rc;
  const size_t smem = (size_t)(3 * p.kblocks + 2) * OPB + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    SRFRD_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  int grid = p.n_tiles * heads;
  if (grid > num_sms()) grid = num_sms();
  attn_fwd_tc_kernel<<<grid, ATC_THREADS, smem, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_attention_bwd(const void* dout, int lddo, const void* q, int ldq, const void* k, const void* v,
                                   int ldkv, void* dq, int lddq, void* dk, void* dv, int lddkv, int64_t B, int L,
                                   int H, int heads, float drop_p, uint64_t seed, uint32_t stream_id,
                                   const float* drop_step, void* stream) {
  SRFRD_REQUIRE(dout && q && k && v && dq && dk && dv, "attention_bwd: null pointer");
  if (B == 0 || L == 0) return 0;
  if (!tc_supported(L, H, heads, ldq, ldkv) || lddo % 8 || lddq % 8 || lddkv % 8)
    return attn_bwd_simt(dout, lddo, q, ldq, k, v, ldkv, dq, lddq, dk, dv, lddkv, B, L, H, heads, drop_p, seed,
                         stream_id, drop_step, stream);
  AttnTc p = {};
  if (int rc = fill(p, B, L, H, heads, drop_p, seed, stream_id, drop_step)) return rc;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.lddq = lddq; p.lddkv = lddkv;
  CUtensorMap tmQ, tmK, tmV, tmDO;
  if (int rc = make_tmap_bf16_2d(&tmQ, q, p.T, H, ldq, TILE, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmK, k, p.T, H, ldkv, TILE, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmV, v, p.T, H, ldkv, TILE, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmDO, dout, p.T, H, lddo, TILE, 64)) return rc;
  const size_t smem = (size_t)(4 * p.kblocks + 4) * OPB + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    SRFRD_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  int grid = p.n_tiles * heads;
  if (grid > num_sms()) grid = num_sms();
  attn_bwd_tc_kernel<<<grid, ATC_THREADS, smem, (cudaStream_t)stream>>>(tmQ, tmK, tmV, tmDO, p);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

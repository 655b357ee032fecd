// Causal self-attention on tcgen05 for 128 < maxlen <= 256 and any head width that is a multiple of 16 (C4: maxlen 200,
// hd = 272) -- SURVEY.md 2.3 row k2; SRFR_model.py:112.
//
// A sequence is two query tiles of 128 rows; query tile i attends key tiles 0..i.  Every operand moves through ONE ring of
// [128 rows x 64 columns] bf16 TMA tiles (16 KB, 128-byte swizzle) that serve as K-major operands (contraction over the 64
// columns: S = Q K^T, dP = dO V^T) or as MN-major operands (contraction over the 128 rows: P V, dS K, dS^T Q, P^T dO) --
// only the descriptor changes.  Head widths above 128 never need more than 64 accumulator columns per MMA: outputs are
// produced in 64-column blocks.
//
//  forward : S[128 x 256] in TMEM (both key tiles), 16 softmax warps (4 per TMEM lane quarter, <= 2 chunks of 32 key
//            columns each, row max / sum exchanged through shared memory), P -> smem, O = P V in 64-column blocks through
//            a ring of 4 TMEM accumulators; per row {max * scale * log2(e), 1 / sum} saved for the backward.
//  backward: FlashAttention-2 decomposition.  With the saved row statistics and delta_i = sum_c dO[i,c] O[i,c] every
//            (query tile, key tile) pair is independent:  P = exp2(S * c - mb_i) / sum_i,  dS = P (dP - delta_i) * scale.
//            Three passes over the pairs, one per output (272 accumulator columns + 128 for S / dP fit TMEM; two outputs
//            do not):  dQ_i += dS K_j  (loop j <= i),   dK_j += dS^T Q_i  and  dV_j += P^T dO_i  (loop i >= j).
//            S and dP share the same 128 TMEM columns: the softmax warps keep P in registers while dP is computed.
#include <stdlib.h>

#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

static constexpr int LT = 128;                 // tile rows
static constexpr int LTB = LT * 64 * 2;        // one [128 x 64] bf16 tile, 16 KB

struct AttnLong {
  int64_t T;
  int B, L, hd, heads, kblocks, nq;
  int ring;                                    // tiles in the operand ring
  float scale, scale_log2e;
  float4* stats;                               // (T * heads): {max * scale * log2e, 1 / sum, delta, -}
  bf16* o; int ldo;                            // forward output
  const bf16* dout; int lddo;                  // backward: delta pre-pass
  bf16* dx; int lddx;                          // backward output of this pass (dq, dk or dv)
  uint64_t drop_seed; uint32_t drop_thresh, drop_stream; float drop_scale; const float* drop_step;
  // Dead query tiles (srfrd_set_attention_live; hybrid packed layout only, where nothing reads a pad query's o / dq):
  // q_lo[b] = first query tile of sequence b that holds a kept token.  Query tiles below it are all dropped padding: their
  // outputs are dead and their dO is zero, so (query tile, key tile) pairs with a dead query tile contribute exactly
  // nothing.  The forward and the dQ pass walk `items` (the n_live[0] live (sequence, head, query tile) items, in order)
  // instead of all of them; the dK / dV passes start each key tile's pair loop at max(key tile, q_lo[b]).
  const int* q_lo; const int* items; const int* n_live;
  const int* tok_row;   // (B * L) packed row of each dense position, -1 = dropped pad slot (its dO is zero): delta pre-pass
};

__device__ __forceinline__ uint32_t lsw128(int r, int c) {     // byte offset of (row r, column c) in a swizzled tile
  return (uint32_t)(r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1)));
}
__device__ __forceinline__ float lex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void lbar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ float ldrop(const AttnLong& p, int64_t bh, int l, int key) {
  const uint64_t idx = ((uint64_t)bh * p.L + l) * p.L + key;
  return dropout_keep(p.drop_seed, p.drop_stream, idx, p.drop_thresh) ? p.drop_scale : 0.f;
}
// 32 probabilities of row r, key chunk c (32 keys) -> 4 x 16 B into the swizzled K-major tile pair `base`
__device__ __forceinline__ void lstore32(uint8_t* base, int r, int c, const uint32_t (&v)[32]) {
  uint8_t* blk = base + (c >> 1) * LTB;
  const int col0 = (c & 1) * 32;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    *reinterpret_cast<uint4*>(blk + lsw128(r, col0 + 8 * u)) =
        make_uint4(pack_bf16x2(__uint_as_float(v[8 * u]), __uint_as_float(v[8 * u + 1])),
                   pack_bf16x2(__uint_as_float(v[8 * u + 2]), __uint_as_float(v[8 * u + 3])),
                   pack_bf16x2(__uint_as_float(v[8 * u + 4]), __uint_as_float(v[8 * u + 5])),
                   pack_bf16x2(__uint_as_float(v[8 * u + 6]), __uint_as_float(v[8 * u + 7])));
}
__device__ __forceinline__ void lzero32(uint8_t* base, int r, int c) {
  uint8_t* blk = base + (c >> 1) * LTB;
#pragma unroll
  for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(blk + lsw128(r, (c & 1) * 32 + 8 * u)) = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------------------------------------------------ forward
static constexpr int ALF_THREADS = 576;        // 16 softmax / epilogue warps + TMA warp + MMA warp

template <bool DROP>
__global__ void __launch_bounds__(ALF_THREADS, 1)
attn_long_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, AttnLong p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* Ps = ring + p.ring * LTB;                       // P: 4 key blocks of 64
  float* xmax = reinterpret_cast<float*>(Ps + 4 * LTB);    // [4 parts][128 rows]
  float* xsum = xmax + 4 * LT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xsum + 4 * LT);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.ring;
  uint64_t* s_full = bars + 2 * p.ring;
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_done = s_full + 2;
  uint64_t* o_full = s_full + 3;                           // [4] accumulator slots
  uint64_t* o_empty = s_full + 7;                          // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 11);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.items ? __ldg(p.n_live) : p.B * p.heads * p.nq;
#define LIVE_ITEM(k) (p.items ? __ldg(p.items + (k)) : (k))

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int i = 0; i < p.ring; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 16); mbar_init(pv_done, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 17) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 256;     // S: 2 key tiles x 128 columns; O: 4 slots x 64 columns
  pdl_prologue_done();
  if (DROP) p.drop_seed = mix_seed(p.drop_seed, p.drop_step);

  if (warp == 16) {
    // ---------------------------------------------------------------- TMA producer: tiles in consumption order
    int slot = 0; uint32_t ph = 0;
    auto push = [&](const CUtensorMap* tm, int col, int row) {
      mbar_wait(&empty[slot], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full[slot], LTB);
        tma_load_2d(ring + slot * LTB, tm, &full[slot], col, row, SRFRD_EVICT_FIRST);
      }
      __syncwarp();
      if (++slot == p.ring) { slot = 0; ph ^= 1; }
    };
    for (int k = blockIdx.x; k < n_items; k += gridDim.x) {
      const int it = LIVE_ITEM(k);
      const int bh = it / p.nq, i = it % p.nq, b = bh / p.heads, h = bh % p.heads;
      const int seq0 = b * p.L, col0 = h * p.hd;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        push(&tmQ, col0 + kb * 64, seq0 + i * LT);
        for (int j = 0; j <= i; ++j) push(&tmK, col0 + kb * 64, seq0 + j * LT);
      }
      for (int cb = 0; cb < p.kblocks; ++cb)
        for (int j = 0; j <= i; ++j) push(&tmV, col0 + cb * 64, seq0 + j * LT);
    }
  } else if (warp == 17) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idS = umma_idesc_bf16(LT, LT, 0, 0);
    const uint64_t dR = umma_smem_desc(smem_u32(ring), 0, 1024);          // K-major view of a ring tile
    const uint64_t mR = umma_smem_desc(smem_u32(ring), LTB, 1024);        // MN-major view (rows = K)
    const uint64_t dP = umma_smem_desc(smem_u32(Ps), 0, 1024);
    int slot = 0; uint32_t ph = 0;
    uint32_t iph = 0;                                                     // per-item phase (s_full, p_full)
    uint32_t ocnt = 0;                                                    // accumulator slot uses so far
    for (int k = blockIdx.x; k < n_items; k += gridDim.x, iph ^= 1) {
      const int i = LIVE_ITEM(k) % p.nq;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const int ksteps = min(4, (p.hd - kb * 64) / 16);
        mbar_wait(&full[slot], ph);
        const int qslot = slot;
        if (++slot == p.ring) { slot = 0; ph ^= 1; }
        for (int j = 0; j <= i; ++j) {
          mbar_wait(&full[slot], ph);
          tc_fence_after();
          if (elect_one()) {
            for (int k = 0; k < ksteps; ++k)
              umma_bf16(tS + j * LT, dR + (uint64_t)(qslot * (LTB >> 4) + 2 * k), dR + (uint64_t)(slot * (LTB >> 4) + 2 * k), idS,
                        (kb | k) != 0);
            umma_commit(&empty[slot]);
            if (j == i) umma_commit(&empty[qslot]);
          }
          __syncwarp();
          if (++slot == p.ring) { slot = 0; ph ^= 1; }
        }
      }
      if (elect_one()) umma_commit(s_full);
      __syncwarp();
      mbar_wait(p_full, iph);                       // P of this item is in shared memory
      tc_fence_after();
      for (int cb = 0; cb < p.kblocks; ++cb, ++ocnt) {
        const int os = ocnt & 3;
        const int ncols = min(64, p.hd - cb * 64);
        const uint32_t idO = umma_idesc_bf16(LT, ncols, 0, 1);
        mbar_wait(&o_empty[os], ((ocnt >> 2) & 1) ^ 1);
        tc_fence_after();
        for (int j = 0; j <= i; ++j) {
          mbar_wait(&full[slot], ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)          // 128 keys of tile j; V tile read MN-major: 16 rows = 2048 B
              umma_bf16(tO + os * 64, dP + (uint64_t)((2 * j + (ks >> 2)) * (LTB >> 4) + (ks & 3) * 2),
                        mR + (uint64_t)(slot * (LTB >> 4) + ks * 128), idO, (j | ks) != 0);
            umma_commit(&empty[slot]);
          }
          __syncwarp();
          if (++slot == p.ring) { slot = 0; ph ^= 1; }
        }
        if (elect_one()) {
          umma_commit(&o_full[os]);
          if (cb == p.kblocks - 1) umma_commit(pv_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + O epilogue
    const int quarter = warp & 3, part = warp >> 2;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int r = quarter * 32 + lane;
    uint32_t iph = 0, ocnt = 0;
    for (int k = blockIdx.x; k < n_items; k += gridDim.x, iph ^= 1) {
      const int it = LIVE_ITEM(k);
      const int bh = it / p.nq, i = it % p.nq, b = bh / p.heads, h = bh % p.heads;
      const int l = i * LT + r;                      // position inside the sequence
      const bool own = l < p.L;
      const int64_t t = (int64_t)b * p.L + l;
      const int jhi = own ? l : -1;                  // last key this row attends (causal, same sequence)
      const int c_hi = 4 * i + quarter;              // last 32-key chunk any row of this quarter needs
      const int c0 = part, c1 = part + 4;            // this warp's chunks
      mbar_wait(s_full, iph);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld32(tS + lane_off + min(c0, c_hi) * 32, s0);
      tmem_ld32(tS + lane_off + min(c1, c_hi) * 32, s1);
      tmem_ld_wait();
      const int k0 = (c0 <= c_hi) ? c0 * 32 : (1 << 20), k1 = (c1 <= c_hi) ? c1 * 32 : (1 << 20);
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        m = (k0 + j <= jhi) ? fmaxf(m, __uint_as_float(s0[j])) : m;
        m = (k1 + j <= jhi) ? fmaxf(m, __uint_as_float(s1[j])) : m;
      }
      xmax[part * LT + r] = m;
      lbar(1 + quarter, 128);                        // the four warps of this lane quarter
#pragma unroll
      for (int q = 0; q < 4; ++q) m = fmaxf(m, xmax[q * LT + r]);
      const float mb = m * p.scale_log2e;
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float e0 = lex2(fmaf(__uint_as_float(s0[j]), p.scale_log2e, -mb));
        float e1 = lex2(fmaf(__uint_as_float(s1[j]), p.scale_log2e, -mb));
        e0 = (k0 + j <= jhi) ? e0 : 0.f;
        e1 = (k1 + j <= jhi) ? e1 : 0.f;
        sum += e0 + e1;
        if (DROP) {
          if (e0 != 0.f) e0 *= ldrop(p, bh, l, k0 + j);
          if (e1 != 0.f) e1 *= ldrop(p, bh, l, k1 + j);
        }
        s0[j] = __float_as_uint(e0); s1[j] = __float_as_uint(e1);
      }
      xsum[part * LT + r] = sum;
      if (k != (int)blockIdx.x) mbar_wait(pv_done, iph ^ 1);      // the previous item's P V MMAs no longer read P
      if (c0 <= c_hi) lstore32(Ps, r, c0, s0);
      if (c1 <= c_hi) lstore32(Ps, r, c1, s1);
      for (int c = part; c < 4 * (i + 1); c += 4)                   // chunks of the attended key tiles beyond the window
        if (c > c_hi) lzero32(Ps, r, c);
      tc_fence_before();
      fence_proxy_async();                           // generic-proxy smem writes -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      lbar(1 + quarter, 128);                        // partial sums of the quarter are visible
      const float tot = xsum[r] + xsum[LT + r] + xsum[2 * LT + r] + xsum[3 * LT + r];
      const float inv = own ? 1.f / tot : 0.f;
      if (part == 0 && own && p.stats) p.stats[t * p.heads + h] = make_float4(mb, inv, 0.f, 0.f);
      // O blocks: 64 columns each, block cb handled by the warps with part == slot of the block
      for (int cb = 0; cb < p.kblocks; ++cb) {
        const uint32_t oc = ocnt + cb;
        if ((int)(oc & 3) != part) continue;
        const int ncols = min(64, p.hd - cb * 64);
        mbar_wait(&o_full[oc & 3], (oc >> 2) & 1);
        tc_fence_after();
        uint32_t a0[32], a1[32];
        tmem_ld32(tO + lane_off + (oc & 3) * 64, a0);
        tmem_ld32(tO + lane_off + (oc & 3) * 64 + 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[oc & 3]);
        if (own) {
          bf16* dst = p.o + t * p.ldo + h * p.hd + cb * 64;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (8 * u >= ncols) break;
            const uint32_t* src = u < 4 ? &a0[8 * u] : &a1[8 * (u - 4)];
            *reinterpret_cast<uint4*>(dst + 8 * u) =
                make_uint4(pack_bf16x2(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv),
                           pack_bf16x2(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv),
                           pack_bf16x2(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv),
                           pack_bf16x2(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv));
          }
        }
      }
      ocnt += p.kblocks;
    }
  }
#undef LIVE_ITEM
  tc_fence_before();
  __syncthreads();
  if (warp == 17) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ----------------------------------------------------------------------------------------------------------- backward
// delta_i = sum_c dO[i, c] * O[i, c] per (token, head): one warp per row, written into stats[...].z
__global__ void __launch_bounds__(256) attn_long_delta_kernel(const bf16* dout, int lddo, const bf16* o, int ldo, float4* stats,
                                                              int64_t T, int heads, int hd, const int* tok_row) {
  pdl_prologue_done();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T * heads) return;
  const int64_t t = row / heads; const int h = (int)(row % heads);
  if (tok_row && __ldg(tok_row + t) < 0) {      // hybrid layout: a dropped pad slot has dO = 0 (srfrd_unpack_rows), so delta = 0
    if (lane == 0) reinterpret_cast<float*>(stats + row)[2] = 0.f;      // without reading 2 x hd values (64 % of C4's rows)
    return;
  }
  const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(dout + t * lddo + h * hd);
  const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(o + t * ldo + h * hd);
  float acc = 0.f;
  for (int c = lane; c < hd / 2; c += 32) {
    const float2 x = __bfloat1622float2(a[c]), y = __bfloat1622float2(b[c]);
    acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
  }
  acc = warp_sum(acc);
  if (lane == 0) reinterpret_cast<float*>(stats + row)[2] = acc;
}

static constexpr int ALB_THREADS = 320;        // 8 softmax / epilogue warps + TMA warp + MMA warp
enum { MODE_DQ = 0, MODE_DK = 1, MODE_DV = 2 };

template <int MODE, bool DROP>
__global__ void __launch_bounds__(ALB_THREADS, 1)
attn_long_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, AttnLong p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* Xs = ring + p.ring * LTB;                       // P (dV pass) or dS (dQ / dK passes): 2 key blocks of 64
  uint64_t* bars = reinterpret_cast<uint64_t*>(Xs + 2 * LTB);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.ring;
  uint64_t* s_full = bars + 2 * p.ring;
  uint64_t *s_read = s_full + 1, *dp_full = s_full + 2, *sd_free = s_full + 3, *x_full = s_full + 4, *x_done = s_full + 5,
           *acc_full = s_full + 6, *acc_free = s_full + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool listed = MODE == MODE_DQ && p.items != nullptr;                // dQ pass: live query tiles only
  const int n_items = listed ? __ldg(p.n_live) : p.B * p.heads * p.nq;
#define LIVE_ITEM(k) (listed ? __ldg(p.items + (k)) : (k))

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    for (int i = 0; i < p.ring; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_read, 8); mbar_init(dp_full, 1); mbar_init(sd_free, 8); mbar_init(x_full, 8);
    mbar_init(x_done, 1); mbar_init(acc_full, 1); mbar_init(acc_free, 8);
    fence_barrier_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tSD = tmem_base, tACC = tmem_base + 128;  // S / dP: 128 columns; output: kblocks x 64 columns
  pdl_prologue_done();
  if (DROP) p.drop_seed = mix_seed(p.drop_seed, p.drop_step);

  // pairs of an item: dQ pass: fixed query tile t, key tiles 0..t;  dK / dV passes: fixed key tile t, query tiles t..nq-1
  // (dK / dV with q_lo: the pairs of key tile t start at its first LIVE query tile)
#define PAIR_LOOP_BEGIN                                                                             \
  const int it = LIVE_ITEM(k);                                                                      \
  const int bh = it / p.nq, tt = it % p.nq, b = bh / p.heads, h = bh % p.heads;                     \
  const int seq0 = b * p.L, col0 = h * p.hd;                                                        \
  const int q0 = (MODE != MODE_DQ && p.q_lo) ? max(tt, __ldg(p.q_lo + b)) : tt;                     \
  const int np_item = (MODE == MODE_DQ) ? tt + 1 : p.nq - q0;                                       \
  for (int pr = 0; pr < np_item; ++pr) {                                                            \
    const int qi = (MODE == MODE_DQ) ? tt : q0 + pr, kj = (MODE == MODE_DQ) ? pr : tt;
#define PAIR_LOOP_END }

  if (warp == 8) {
    // ---------------------------------------------------------------- TMA producer
    int slot = 0; uint32_t ph = 0;
    auto push = [&](const CUtensorMap* tm, int col, int row) {
      mbar_wait(&empty[slot], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full[slot], LTB);
        tma_load_2d(ring + slot * LTB, tm, &full[slot], col, row, SRFRD_EVICT_FIRST);
      }
      __syncwarp();
      if (++slot == p.ring) { slot = 0; ph ^= 1; }
    };
    for (int k = blockIdx.x; k < n_items; k += gridDim.x) {
      PAIR_LOOP_BEGIN
        for (int kb = 0; kb < p.kblocks; ++kb) {
          push(&tmQ, col0 + kb * 64, seq0 + qi * LT);
          push(&tmK, col0 + kb * 64, seq0 + kj * LT);
        }
        if (MODE != MODE_DV)
          for (int kb = 0; kb < p.kblocks; ++kb) {
            push(&tmDO, col0 + kb * 64, seq0 + qi * LT);
            push(&tmV, col0 + kb * 64, seq0 + kj * LT);
          }
        for (int cb = 0; cb < p.kblocks; ++cb) {
          if (MODE == MODE_DQ) push(&tmK, col0 + cb * 64, seq0 + kj * LT);
          else if (MODE == MODE_DK) push(&tmQ, col0 + cb * 64, seq0 + qi * LT);
          else push(&tmDO, col0 + cb * 64, seq0 + qi * LT);
        }
      PAIR_LOOP_END
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idKK = umma_idesc_bf16(LT, LT, 0, 0);
    const uint64_t dR = umma_smem_desc(smem_u32(ring), 0, 1024);          // K-major view of a ring tile
    const uint64_t mR = umma_smem_desc(smem_u32(ring), LTB, 1024);        // MN-major view (rows = K)
    const uint64_t kX = umma_smem_desc(smem_u32(Xs), 0, 1024);            // X as K-major A (dQ = dS K)
    const uint64_t mX = umma_smem_desc(smem_u32(Xs), LTB, 1024);          // X^T as MN-major A (dK, dV)
    int slot = 0; uint32_t ph = 0;
    uint32_t np = 0, ni = 0;                                              // pairs / items so far
    for (int k = blockIdx.x; k < n_items; k += gridDim.x, ++ni) {
      PAIR_LOOP_BEGIN
        (void)qi; (void)kj; (void)seq0; (void)col0; (void)h;
        const uint32_t pp = np & 1;
        mbar_wait(sd_free, pp ^ 1);                  // previous pair's S / dP have been read out
        tc_fence_after();
        for (int ph2 = 0; ph2 < (MODE == MODE_DV ? 1 : 2); ++ph2) {
          if (ph2 == 1) { mbar_wait(s_read, pp); tc_fence_after(); }      // S is in registers: dP may overwrite it
          for (int kb = 0; kb < p.kblocks; ++kb) {
            const int ksteps = min(4, (p.hd - kb * 64) / 16);
            mbar_wait(&full[slot], ph);
            const int aslot = slot;
            if (++slot == p.ring) { slot = 0; ph ^= 1; }
            mbar_wait(&full[slot], ph);
            tc_fence_after();
            if (elect_one()) {
              for (int k = 0; k < ksteps; ++k)
                umma_bf16(tSD, dR + (uint64_t)(aslot * (LTB >> 4) + 2 * k), dR + (uint64_t)(slot * (LTB >> 4) + 2 * k), idKK,
                          (kb | k) != 0);
              umma_commit(&empty[aslot]);
              umma_commit(&empty[slot]);
            }
            __syncwarp();
            if (++slot == p.ring) { slot = 0; ph ^= 1; }
          }
          if (elect_one()) umma_commit(ph2 == 0 ? s_full : dp_full);
          __syncwarp();
        }
        mbar_wait(x_full, pp);                       // P / dS of this pair is in shared memory
        if (pr == 0) mbar_wait(acc_free, (ni & 1) ^ 1);                   // previous item's output has been read out
        tc_fence_after();
        for (int cb = 0; cb < p.kblocks; ++cb) {
          const int ncols = min(64, p.hd - cb * 64);
          mbar_wait(&full[slot], ph);
          tc_fence_after();
          if (elect_one()) {
            if (MODE == MODE_DQ) {
              const uint32_t id = umma_idesc_bf16(LT, ncols, 0, 1);       // A = dS K-major, B = K tile MN-major
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                umma_bf16(tACC + cb * 64, kX + (uint64_t)((ks >> 2) * (LTB >> 4) + (ks & 3) * 2),
                          mR + (uint64_t)(slot * (LTB >> 4) + ks * 128), id, (pr | ks) != 0);
            } else {
              const uint32_t id = umma_idesc_bf16(LT, ncols, 1, 1);       // A = X^T MN-major, B = Q / dO tile MN-major
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                umma_bf16(tACC + cb * 64, mX + (uint64_t)(ks * 128), mR + (uint64_t)(slot * (LTB >> 4) + ks * 128), id,
                          (pr | ks) != 0);
            }
            umma_commit(&empty[slot]);
            if (cb == p.kblocks - 1) {
              umma_commit(x_done);
              if (pr == np_item - 1) umma_commit(acc_full);
            }
          }
          __syncwarp();
          if (++slot == p.ring) { slot = 0; ph ^= 1; }
        }
        ++np;
      PAIR_LOOP_END
    }
  } else {
    // ---------------------------------------------------------------- softmax backward + output epilogue
    const int quarter = warp & 3, part = warp >> 2;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int r = quarter * 32 + lane;
    const int c0 = part, c1 = part + 2;              // this warp's 32-key chunks of the pair
    uint32_t np = 0, ni = 0;
    for (int k = blockIdx.x; k < n_items; k += gridDim.x, ++ni) {
      PAIR_LOOP_BEGIN
        const uint32_t pp = np & 1;
        const int l = qi * LT + r;                   // query position inside the sequence
        const bool own = l < p.L;
        const int64_t t = (int64_t)seq0 + l;
        float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
        if (own) st = p.stats[t * p.heads + h];
        const int kmax = own ? (qi == kj ? r : LT - 1) : -1;              // last key column of this pair the row attends
        mbar_wait(s_full, pp);
        tc_fence_after();
        uint32_t s0[32], s1[32];
        tmem_ld32(tSD + lane_off + c0 * 32, s0);
        tmem_ld32(tSD + lane_off + c1 * 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(MODE == MODE_DV ? sd_free : s_read);
#pragma unroll
        for (int j = 0; j < 32; ++j) {               // P (undropped), kept in registers
          float e0 = lex2(fmaf(__uint_as_float(s0[j]), p.scale_log2e, -st.x)) * st.y;
          float e1 = lex2(fmaf(__uint_as_float(s1[j]), p.scale_log2e, -st.x)) * st.y;
          s0[j] = __float_as_uint((c0 * 32 + j <= kmax) ? e0 : 0.f);
          s1[j] = __float_as_uint((c1 * 32 + j <= kmax) ? e1 : 0.f);
        }
        if (MODE == MODE_DV) {
          if (DROP) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (s0[j]) s0[j] = __float_as_uint(__uint_as_float(s0[j]) * ldrop(p, bh, l, kj * LT + c0 * 32 + j));
              if (s1[j]) s1[j] = __float_as_uint(__uint_as_float(s1[j]) * ldrop(p, bh, l, kj * LT + c1 * 32 + j));
            }
          }
        } else {
          mbar_wait(dp_full, pp);
          tc_fence_after();
          uint32_t g0[32], g1[32];
          tmem_ld32(tSD + lane_off + c0 * 32, g0);
          tmem_ld32(tSD + lane_off + c1 * 32, g1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sd_free);
#pragma unroll
          for (int j = 0; j < 32; ++j) {             // dS = P (dP * mask - delta) * scale
            float d0 = __uint_as_float(g0[j]), d1 = __uint_as_float(g1[j]);
            if (DROP) {
              if (s0[j]) d0 *= ldrop(p, bh, l, kj * LT + c0 * 32 + j);
              if (s1[j]) d1 *= ldrop(p, bh, l, kj * LT + c1 * 32 + j);
            }
            s0[j] = __float_as_uint(__uint_as_float(s0[j]) * (d0 - st.z) * p.scale);
            s1[j] = __float_as_uint(__uint_as_float(s1[j]) * (d1 - st.z) * p.scale);
          }
        }
        mbar_wait(x_done, pp ^ 1);                   // previous pair's output MMAs no longer read X
        lstore32(Xs, r, c0, s0);
        lstore32(Xs, r, c1, s1);
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(x_full);
        ++np;
      PAIR_LOOP_END
      // ---- output rows of tile tt: 64-column blocks alternate between the two warps of a quarter
      mbar_wait(acc_full, ni & 1);
      tc_fence_after();
      {
        const int it = LIVE_ITEM(k);
        const int bh2 = it / p.nq, tt2 = it % p.nq, b2 = bh2 / p.heads, h2 = bh2 % p.heads;
        const int lo = tt2 * LT + r;
        const bool oown = lo < p.L;
        bf16* dst = p.dx + ((int64_t)b2 * p.L + lo) * p.lddx + h2 * p.hd;
        for (int cb = part; cb < p.kblocks; cb += 2) {
          const int ncols = min(64, p.hd - cb * 64);
          uint32_t a0[32], a1[32];
          tmem_ld32(tACC + lane_off + cb * 64, a0);
          tmem_ld32(tACC + lane_off + cb * 64 + 32, a1);
          tmem_ld_wait();
          if (oown) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (8 * u >= ncols) break;
              const uint32_t* src = u < 4 ? &a0[8 * u] : &a1[8 * (u - 4)];
              *reinterpret_cast<uint4*>(dst + cb * 64 + 8 * u) =
                  make_uint4(pack_bf16x2(__uint_as_float(src[0]), __uint_as_float(src[1])),
                             pack_bf16x2(__uint_as_float(src[2]), __uint_as_float(src[3])),
                             pack_bf16x2(__uint_as_float(src[4]), __uint_as_float(src[5])),
                             pack_bf16x2(__uint_as_float(src[6]), __uint_as_float(src[7])));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
    }
  }
#undef PAIR_LOOP_BEGIN
#undef PAIR_LOOP_END
#undef LIVE_ITEM
  tc_fence_before();
  __syncthreads();
  if (warp == 9) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- live query tiles of the hybrid packed layout (see AttnLong)
static thread_local const int* g_live_qlo = nullptr;
static thread_local const int* g_live_items = nullptr;
static thread_local const int* g_live_n = nullptr;
static thread_local const int* g_live_tok_row = nullptr;
void attn_long_set_live(const int* q_lo, const int* items, const int* n_live, const int* tok_row) {
  g_live_qlo = q_lo; g_live_items = items; g_live_n = n_live; g_live_tok_row = tok_row;
}

// q_lo[b] = tile of the first position of sequence b whose token is kept (tok_row >= 0), at most nq - 1; one warp a sequence
__global__ void __launch_bounds__(256) attn_qlo_kernel(const int* tok_row, int64_t B, int L, int nq, int* q_lo) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int first = L;
  for (int l0 = 0; l0 < L && first == L; l0 += 32) {
    const int l = l0 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, l < L && __ldg(tok_row + b * L + l) >= 0);
    if (m) first = l0 + __ffs(m) - 1;
  }
  if (lane == 0) q_lo[b] = min(first / LT, nq - 1);
}
// items[] = the live (sequence, head, query tile) items in increasing order, n_live[0] = their number.  ONE block.
__global__ void __launch_bounds__(1024) attn_items_kernel(const int* q_lo, int64_t B, int heads, int nq, int* items, int* n_live) {
  __shared__ int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (int)((B + 1023) / 1024);
  const int64_t b0 = (int64_t)tid * per, b1 = min(B, b0 + per);
  int local = 0;
  for (int64_t b = b0; b < b1; ++b) local += heads * (nq - q_lo[b]);
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    wsum[lane] = v;
    if (lane == 31) n_live[0] = v;
  }
  __syncthreads();
  int at = incl - local + (warp ? wsum[warp - 1] : 0);
  for (int64_t b = b0; b < b1; ++b) {
    const int lo = q_lo[b];
    for (int h = 0; h < heads; ++h)
      for (int t = lo; t < nq; ++t) items[at++] = (int)((b * heads + h) * nq + t);
  }
}

static int fill_long(AttnLong& p, int64_t B, int L, int H, int heads, float drop_p, uint64_t seed, uint32_t stream_id,
                     const float* drop_step) {
  p.q_lo = g_live_qlo; p.items = g_live_items; p.n_live = g_live_n; p.tok_row = g_live_tok_row;
  p.B = (int)B; p.T = B * L; p.L = L; p.heads = heads; p.hd = H / heads;
  p.kblocks = (p.hd + 63) / 64;
  p.nq = (L + LT - 1) / LT;
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

bool attn_long_supported(int64_t B, int L, int H, int heads, int ldq, int ldkv) {
  if (heads <= 0 || H % heads) return false;
  const int hd = H / heads;
  return L > LT && L <= 2 * LT && hd % 16 == 0 && hd <= 512 && ldq % 8 == 0 && ldkv % 8 == 0 && B * L < (1ll << 31);
}

int attn_long_fwd(const void* q, int ldq, const void* k, const void* v, int ldkv, void* o, int ldo, float* stats, int64_t B,
                  int L, int H, int heads, float drop_p, uint64_t seed, uint32_t stream_id, const float* drop_step,
                  void* stream) {
  SRFRD_REQUIRE(ldo % 8 == 0 && (((uintptr_t)o) & 15) == 0, "attention_fwd: output must be 16-byte aligned with ld %% 8 == 0");
  AttnLong p = {};
  if (int rc = fill_long(p, B, L, H, heads, drop_p, seed, stream_id, drop_step)) return rc;
  p.o = (bf16*)o; p.ldo = ldo; p.stats = reinterpret_cast<float4*>(stats);
  CUtensorMap tmQ, tmK, tmV;
  if (int rc = make_tmap_bf16_2d(&tmQ, q, p.T, H, ldq, LT, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmK, k, p.T, H, ldkv, LT, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmV, v, p.T, H, ldkv, LT, 64)) return rc;
  p.ring = 9;
  const size_t smem = (size_t)(p.ring + 4) * LTB + 8 * LT * sizeof(float) + 1024 + 512;
  static bool attr = false;
  if (!attr) {
    SRFRD_CUDA(cudaFuncSetAttribute(attn_long_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SRFRD_CUDA(cudaFuncSetAttribute(attn_long_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  int grid = p.B * heads * p.nq;
  if (grid > num_sms()) grid = num_sms();
  if (p.drop_thresh)
    SRFRD_CUDA(launch_pdl(attn_long_fwd_kernel<true>, dim3(grid), dim3(ALF_THREADS), smem, (cudaStream_t)stream, tmQ, tmK, tmV, p));
  else
    SRFRD_CUDA(launch_pdl(attn_long_fwd_kernel<false>, dim3(grid), dim3(ALF_THREADS), smem, (cudaStream_t)stream, tmQ, tmK, tmV, p));
  return 0;
}

template <int MODE>
static int launch_long_bwd(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmDO,
                           AttnLong& p, cudaStream_t stream) {
  p.ring = 11;
  const size_t smem = (size_t)(p.ring + 2) * LTB + 1024 + 512;
  static bool attr = false;
  if (!attr) {
    SRFRD_CUDA(cudaFuncSetAttribute(attn_long_bwd_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SRFRD_CUDA(cudaFuncSetAttribute(attn_long_bwd_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  int grid = p.B * p.heads * p.nq;
  if (grid > num_sms()) grid = num_sms();
  if (p.drop_thresh)
    SRFRD_CUDA(launch_pdl(attn_long_bwd_kernel<MODE, true>, dim3(grid), dim3(ALB_THREADS), smem, stream, tmQ, tmK, tmV, tmDO, p));
  else
    SRFRD_CUDA(launch_pdl(attn_long_bwd_kernel<MODE, false>, dim3(grid), dim3(ALB_THREADS), smem, stream, tmQ, tmK, tmV, tmDO, p));
  return 0;
}

int attn_long_bwd(const void* dout, int lddo, const void* q, int ldq, const void* k, const void* v, int ldkv, const void* o,
                  int ldo, float* stats, void* dq, int lddq, void* dk, void* dv, int lddkv, int64_t B, int L, int H, int heads,
                  float drop_p, uint64_t seed, uint32_t stream_id, const float* drop_step, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRFRD_REQUIRE(o && stats, "attention_bwd: maxlen > 128 needs the forward output and row statistics (o, stats)");
  SRFRD_REQUIRE(lddq % 8 == 0 && lddkv % 8 == 0 && lddo % 8 == 0 && ldo % 2 == 0 &&
                    ((((uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv)) & 15) == 0,
                "attention_bwd: gradients must be 16-byte aligned with leading dimensions %% 8 == 0");
  AttnLong p = {};
  if (int rc = fill_long(p, B, L, H, heads, drop_p, seed, stream_id, drop_step)) return rc;
  p.stats = reinterpret_cast<float4*>(stats);
  CUtensorMap tmQ, tmK, tmV, tmDO;
  if (int rc = make_tmap_bf16_2d(&tmQ, q, p.T, H, ldq, LT, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmK, k, p.T, H, ldkv, LT, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmV, v, p.T, H, ldkv, LT, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmDO, dout, p.T, H, lddo, LT, 64)) return rc;
  const int64_t rows = p.T * heads;
  SRFRD_CUDA(launch_pdl(attn_long_delta_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, stream, (const bf16*)dout, lddo,
                        (const bf16*)o, ldo, p.stats, p.T, heads, p.hd, p.tok_row));
  p.dx = (bf16*)dq; p.lddx = lddq;
  if (int rc = launch_long_bwd<MODE_DQ>(tmQ, tmK, tmV, tmDO, p, stream)) return rc;
  p.dx = (bf16*)dk; p.lddx = lddkv;
  if (int rc = launch_long_bwd<MODE_DK>(tmQ, tmK, tmV, tmDO, p, stream)) return rc;
  p.dx = (bf16*)dv; p.lddx = lddkv;
  return launch_long_bwd<MODE_DV>(tmQ, tmK, tmV, tmDO, p, stream);
}

}  // namespace srfrd

extern "C" int srfrd_set_attention_live(const int* q_lo, const int* items, const int* n_live, const int* tok_row) {
  SRFRD_REQUIRE((q_lo && items && n_live && tok_row) || (!q_lo && !items && !n_live && !tok_row),
                "set_attention_live: all four pointers or none");
  srfrd::attn_long_set_live(q_lo, items, n_live, tok_row);
  return 0;
}

extern "C" int srfrd_attention_live_items(const int* tok_row, int64_t B, int L, int heads, int* q_lo, int* items, int* n_live,
                                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRFRD_REQUIRE(tok_row && q_lo && items && n_live, "attention_live_items: null pointer");
  SRFRD_REQUIRE(B >= 1 && B <= 16384 * 64 && L >= 1 && heads >= 1, "attention_live_items: bad shape");
  const int nq = (L + srfrd::LT - 1) / srfrd::LT;
  srfrd::attn_qlo_kernel<<<(unsigned)((B + 7) / 8), 256, 0, stream>>>(tok_row, B, L, nq, q_lo);
  SRFRD_LAUNCH_CHECK();
  srfrd::attn_items_kernel<<<1, 1024, 0, stream>>>(q_lo, B, heads, nq, items, n_live);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

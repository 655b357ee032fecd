// K8: full-catalogue scoring  feats (U, D) x table (N, D)^T  on tcgen05 with a streaming per-row top-10
// in the epilogue; logits never reach HBM.  K9: merge of partial / per-shard candidate lists.
// (SURVEY.md rows A9/A10 + full-catalogue extension: == predict(..., label = arange(1, N+1)) + stable top-k
//  with the tie-break (score desc, item id asc).)
//
// A work unit is (256 users) x (a chunk of the item rows).  The two 128-user tiles (TMEM lanes) stay in smem
// for the whole unit while 128-item tiles stream through a TMA ring and are multiplied against BOTH user
// tiles, which halves the L2 -> SM traffic per score (with one user tile per CTA the kernel was L2-bound at
// ~5 TB/s: every CTA streams the whole table).  Accumulators (2 user tiles x 128 columns) are double-buffered
// in TMEM so the tensor pipe runs ahead of the scan.  Each epilogue thread owns one user row x 128 columns
// and keeps that row's running top-10 in registers: group maxima are compared with the current 10th best
// first, so the insertion path is rare after warm-up.
// n_split = 2/3 feeds hi/lo bf16 splits of the fp32 user features as extra K (near-fp32 scores).
#include <math.h>

#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

static constexpr int TK = 10;               // list length kept per row
static constexpr int TILE_U = 128;          // users per MMA (TMEM lanes)
static constexpr int UBS = 2;               // user blocks per work unit: each item tile is reused for 256 users
static constexpr int GROUP_U = TILE_U * UBS;
static constexpr int TILE_I = 128;          // items per tile (TMEM columns per user block)
static constexpr int KB = 64;               // k-block (bf16 elements)
static constexpr int TOPK_THREADS = 320;    // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (4 per user block)
static constexpr int A_TILE_BYTES = TILE_U * KB * 2;   // 16 KB
static constexpr int B_TILE_BYTES = TILE_I * KB * 2;   // 16 KB
static constexpr int ACC_STRIDE = UBS * TILE_I;        // TMEM columns per accumulator stage (256)

struct TopkShape {
  int U, D;
  int n_split;            // 1..3 user-feature splits (each a (U, D) bf16 matrix stacked along rows: split s at row s*U_pad)
  int u_pad;              // row offset between splits in the feats tensor
  int row_lo, row_hi;     // candidate rows of the (local) table: [row_lo, row_hi)
  int64_t id_base;        // global item id of local row 0
  int kblocks;            // ceil(D / 64)
  int tiles_total;        // item tiles over [row_lo, row_hi)
  int chunks, tiles_per_chunk, ugroups;
  int stages;
  float* out_scores;      // (U, chunks, TK)
  int* out_ids;           // (U, chunks, TK)  global ids (int32), -1 = empty
};

__device__ __forceinline__ void list_insert(float (&ts)[TK], int (&ti)[TK], float s, int id) {
  ts[TK - 1] = s; ti[TK - 1] = id;
#pragma unroll
  for (int r = TK - 1; r > 0; --r) {
    if (ts[r] > ts[r - 1]) {          // strict: an equal score never overtakes an earlier (lower) id
      const float fs = ts[r]; ts[r] = ts[r - 1]; ts[r - 1] = fs;
      const int is = ti[r]; ti[r] = ti[r - 1]; ti[r - 1] = is;
    }
  }
}

__global__ void __launch_bounds__(TOPK_THREADS, 1)
catalogue_topk_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmE, TopkShape s) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_tiles = UBS * s.n_split * s.kblocks;
  const int a_bytes = a_tiles * A_TILE_BYTES;
  uint8_t* smA = smem;                                  // [ub][split][kblock] tiles of 128 users x 64 features
  uint8_t* smB = smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + s.stages * B_TILE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + s.stages;
  uint64_t* tfull = bars + 2 * s.stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* afull = tempty + 2;
  uint64_t* aempty = afull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = s.ugroups * s.chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmF);
    tma_prefetch_desc(&tmE);
    for (int i = 0; i < s.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    mbar_init(afull, 1);
    mbar_init(aempty, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, uphase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int ug = unit / s.chunks, ch = unit % s.chunks;
        const int t0 = ch * s.tiles_per_chunk, t1 = min(s.tiles_total, t0 + s.tiles_per_chunk);
        mbar_wait(aempty, uphase ^ 1);                    // previous unit's MMAs no longer read the user tiles
        mbar_expect_tx(afull, a_bytes);
        for (int ub = 0; ub < UBS; ++ub)
          for (int sp = 0; sp < s.n_split; ++sp)
            for (int kb = 0; kb < s.kblocks; ++kb)
              tma_load_2d(smA + ((ub * s.n_split + sp) * s.kblocks + kb) * A_TILE_BYTES, &tmF, afull, kb * KB,
                          sp * s.u_pad + ug * GROUP_U + ub * TILE_U, SRFRD_EVICT_LAST);
        uphase ^= 1;
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < s.kblocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], B_TILE_BYTES);
            tma_load_2d(smB + stage * B_TILE_BYTES, &tmE, &full[stage], kb * KB, s.row_lo + t * TILE_I,
                        SRFRD_EVICT_NORMAL);
            if (++stage == s.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TILE_U, TILE_I, 0, 0);
      int stage = 0; uint32_t phase = 0, uphase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int ch = unit % s.chunks;
        const int t0 = ch * s.tiles_per_chunk, t1 = min(s.tiles_total, t0 + s.tiles_per_chunk);
        mbar_wait(afull, uphase);
        uphase ^= 1;
        tc_fence_after();
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&tempty[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t tacc = tmem_base + as * ACC_STRIDE;
          for (int kb = 0; kb < s.kblocks; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t b0 = smem_u32(smB + stage * B_TILE_BYTES);
            const int ksteps = min(KB / 16, (s.D - kb * KB + 15) / 16);
            for (int ub = 0; ub < UBS; ++ub)               // the item tile is read from smem once per user block
              for (int sp = 0; sp < s.n_split; ++sp) {
                const uint32_t a0 = smem_u32(smA + ((ub * s.n_split + sp) * s.kblocks + kb) * A_TILE_BYTES);
                for (int k = 0; k < ksteps; ++k)
                  umma_bf16(tacc + ub * TILE_I, umma_smem_desc(a0 + k * 32, 0, 1024),
                            umma_smem_desc(b0 + k * 32, 0, 1024), idesc, (kb | sp | k) != 0);
              }
            umma_commit(&empty[stage]);
            if (++stage == s.stages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull[as]);
          if (++as == 2) { as = 0; aphase ^= 1; }
        }
        umma_commit(aempty);
      }
    }
  } else {
    const int e = warp - 2;
    const int quarter = warp & 3, ub = e >> 2;
    int as = 0; uint32_t aphase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int ug = unit / s.chunks, ch = unit % s.chunks;
      const int t0 = ch * s.tiles_per_chunk, t1 = min(s.tiles_total, t0 + s.tiles_per_chunk);
      const int urow = ug * GROUP_U + ub * TILE_U + quarter * 32 + lane;
      float ts[TK]; int ti[TK];
#pragma unroll
      for (int r = 0; r < TK; ++r) { ts[r] = -INFINITY; ti[r] = -1; }
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&tfull[as], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * ACC_STRIDE + ub * TILE_I;
        const int col_row0 = s.row_lo + t * TILE_I;                  // local table row of column 0
        // all 128 scores of this thread's row in flight at once (one TMEM round trip per tile)
        uint32_t raw[4][32];
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld32(taddr + g * 32, raw[g]);
        tmem_ld_wait();
        float gm[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          // FMNMX3: two scores folded per ALU instruction
          float m0 = fmax3(__uint_as_float(raw[g][0]), __uint_as_float(raw[g][1]), __uint_as_float(raw[g][2]));
          float m1 = fmax3(__uint_as_float(raw[g][3]), __uint_as_float(raw[g][4]), __uint_as_float(raw[g][5]));
#pragma unroll
          for (int j = 6; j < 30; j += 4) {
            m0 = fmax3(m0, __uint_as_float(raw[g][j]), __uint_as_float(raw[g][j + 1]));
            m1 = fmax3(m1, __uint_as_float(raw[g][j + 2]), __uint_as_float(raw[g][j + 3]));
          }
          gm[g] = fmax3(fmaxf(m0, m1), __uint_as_float(raw[g][30]), __uint_as_float(raw[g][31]));
        }
        const float tmax = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
        // Threshold-first: only if some lane of the warp beats its current 10th best do we look closer.
        // Candidate columns are found with a warp-wide OR of per-lane compare masks and re-read from TMEM
        // one column at a time (warp-uniform address), so the insertion code exists once and runs rarely.
        if (__any_sync(0xffffffffu, tmax > ts[TK - 1])) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (!__any_sync(0xffffffffu, gm[g] > ts[TK - 1])) continue;
            uint32_t mask = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(raw[g][j]) > ts[TK - 1]) ? (1u << j) : 0u;
            uint32_t wmask = __reduce_or_sync(0xffffffffu, mask);
            const int base = col_row0 + g * 32;
            while (wmask) {
              const int j = __ffs(wmask) - 1;
              wmask &= wmask - 1;
              uint32_t one;
              tmem_ld1(taddr + g * 32 + j, one);
              tmem_ld_wait();
              const float sc = __uint_as_float(one);
              if (sc > ts[TK - 1] && base + j < s.row_hi) list_insert(ts, ti, sc, base + j);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      if (urow < s.U) {
        const size_t o = ((size_t)urow * s.chunks + ch) * TK;
#pragma unroll
        for (int r = 0; r < TK; ++r) {
          s.out_scores[o + r] = ts[r];
          s.out_ids[o + r] = ti[r] < 0 ? -1 : (int)(s.id_base + ti[r]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// K9: merge `nlists` candidate lists of length TK per user into the best k, order (score desc, id asc).
__global__ void merge_topk_kernel(const float* sc, const int* ids, int64_t U, int nlists, int k, float* out_sc,
                                  int64_t* out_ids) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  float ts[TK]; int ti[TK];
#pragma unroll
  for (int r = 0; r < TK; ++r) { ts[r] = -INFINITY; ti[r] = -1; }
  const float* s0 = sc + (size_t)u * nlists * TK;
  const int* i0 = ids + (size_t)u * nlists * TK;
  for (int n = 0; n < nlists * TK; ++n) {
    const float v = s0[n]; const int id = i0[n];
    if (id < 0) continue;
    const bool better = (ti[TK - 1] < 0) || v > ts[TK - 1] || (v == ts[TK - 1] && id < ti[TK - 1]);
    if (!better) continue;
    ts[TK - 1] = v; ti[TK - 1] = id;
#pragma unroll
    for (int r = TK - 1; r > 0; --r) {
      const bool up = (ti[r - 1] < 0) || ts[r] > ts[r - 1] || (ts[r] == ts[r - 1] && ti[r] < ti[r - 1]);
      if (up) {
        const float fs = ts[r]; ts[r] = ts[r - 1]; ts[r - 1] = fs;
        const int is = ti[r]; ti[r] = ti[r - 1]; ti[r - 1] = is;
      }
    }
  }
  for (int r = 0; r < k; ++r) {
    out_sc[u * k + r] = ts[r];
    out_ids[u * k + r] = ti[r];
  }
}

}  // namespace srfrd

using namespace srfrd;

extern "C" int srfrd_catalogue_topk_plan(int64_t U, int64_t n_rows, int64_t row_lo, int* chunks_out) {
  SRFRD_REQUIRE(chunks_out, "catalogue_topk_plan: null output");
  const int64_t tiles = (n_rows - row_lo + TILE_I - 1) / TILE_I;
  const int64_t ublocks = (U + GROUP_U - 1) / GROUP_U;
  // One wave of work units: every unit restarts its top-10 lists cold (about 10 ln(n/10) insertions per row
  // over n items), so item chunks are only used to occupy SMs that the user blocks alone would leave idle.
  int64_t chunks = ublocks > 0 ? num_sms() / ublocks : 1;
  if (chunks > tiles / 4) chunks = tiles / 4;
  if (chunks < 1) chunks = 1;
  const int64_t per = (tiles + chunks - 1) / chunks;
  chunks = per > 0 ? (tiles + per - 1) / per : 1;
  if (chunks < 1) chunks = 1;
  *chunks_out = (int)chunks;
  return 0;
}

extern "C" int srfrd_catalogue_topk(const void* feats_bf16, int64_t U, int64_t u_pad, int n_split, const void* table_bf16,
                                    int64_t n_rows, int64_t row_lo, int64_t id_base, int D, int ld_feats, int ld_table,
                                    int chunks, float* part_scores, int* part_ids, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRFRD_REQUIRE(feats_bf16 && table_bf16 && part_scores && part_ids, "catalogue_topk: null pointer");
  SRFRD_REQUIRE(n_split >= 1 && n_split <= 3, "catalogue_topk: n_split must be 1..3");
  SRFRD_REQUIRE(D % 16 == 0 && ld_feats % 8 == 0 && ld_table % 8 == 0, "catalogue_topk: D %% 16 and ld %% 8 required (D=%d)", D);
  SRFRD_REQUIRE(U > 0 && n_rows > row_lo && row_lo >= 0, "catalogue_topk: empty problem");
  SRFRD_REQUIRE(n_rows < (1ll << 31) && id_base + n_rows < (1ll << 31), "catalogue_topk: ids must fit int32");
  SRFRD_REQUIRE(chunks >= 1, "catalogue_topk: chunks must be >= 1 (use srfrd_catalogue_topk_plan)");
  TopkShape s;
  s.U = (int)U; s.D = D; s.n_split = n_split; s.u_pad = (int)u_pad;
  s.row_lo = (int)row_lo; s.row_hi = (int)n_rows; s.id_base = id_base;
  s.kblocks = (D + KB - 1) / KB;
  s.tiles_total = (int)((n_rows - row_lo + TILE_I - 1) / TILE_I);
  s.chunks = chunks;
  s.tiles_per_chunk = (s.tiles_total + chunks - 1) / chunks;
  s.ugroups = (int)((U + GROUP_U - 1) / GROUP_U);
  const int a_bytes = UBS * n_split * s.kblocks * A_TILE_BYTES;
  s.stages = (int)((210 * 1024 - a_bytes) / B_TILE_BYTES);
  SRFRD_REQUIRE(s.stages >= 2, "catalogue_topk: D=%d with n_split=%d does not fit shared memory", D, n_split);
  if (s.stages > 8) s.stages = 8;
  s.out_scores = part_scores; s.out_ids = part_ids;
  const size_t smem = (size_t)a_bytes + (size_t)s.stages * B_TILE_BYTES + 1024 + 256;
  CUtensorMap tmF, tmE;
  if (int rc = make_tmap_bf16_2d(&tmF, feats_bf16, (uint64_t)((n_split - 1) * u_pad + U), D, ld_feats, TILE_U, KB)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmE, table_bf16, n_rows, D, ld_table, TILE_I, KB)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    SRFRD_CUDA(cudaFuncSetAttribute(catalogue_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int grid = s.ugroups * s.chunks;
  if (grid > num_sms()) grid = num_sms();
  catalogue_topk_kernel<<<grid, TOPK_THREADS, smem, stream>>>(tmF, tmE, s);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_merge_topk(const float* scores, const int* ids, int64_t U, int nlists, int k, float* out_scores,
                                int64_t* out_ids, void* stream) {
  SRFRD_REQUIRE(scores && ids && out_scores && out_ids, "merge_topk: null pointer");
  SRFRD_REQUIRE(k >= 1 && k <= TK, "merge_topk: k must be in 1..%d", TK);
  if (U == 0) return 0;
  merge_topk_kernel<<<(unsigned)((U + 127) / 128), 128, 0, (cudaStream_t)stream>>>(scores, ids, U, nlists, k, out_scores, out_ids);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

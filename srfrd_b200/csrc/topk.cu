// K8: full-catalogue scoring  feats (U, D) x table (N, D)^T  on tcgen05 with a streaming per-row top-10
// in the epilogue; logits never reach HBM.  K9: merge of partial / per-shard candidate lists.
// (SURVEY.md rows A9/A10 + full-catalogue extension: == predict(..., label = arange(1, N+1)) + stable top-k
//  with the tie-break (score desc, item id asc).)
//
// A work unit is (UBS x 128 users) x (a chunk of the item rows).
//  * The user tiles are STATIONARY IN TENSOR MEMORY: at the start of a unit the epilogue warps copy their rows
//    (bf16, packed two per 32-bit column) into TMEM with tcgen05.st and every MMA takes its A operand from there
//    (tcgen05.mma TS form).  With A in shared memory the SS-form MMA of a 128 x N x 16 step reads 4 KB of A plus
//    N*32 B of B, which saturates the 128 B/clk shared-memory port while TMA is also writing tiles (measured:
//    the MMA pipe sat at 35 % with the epilogue switched off entirely).
//  * Every item tile (NT rows, streamed through a TMA ring) is multiplied against ALL UBS user tiles, which
//    divides the L2 -> SM traffic per score by UBS (one user tile per CTA means every CTA streams the whole
//    table: 16 GB per pass at C3, L2-bound).
//  * Accumulators (UBS x NT columns) are double-buffered in TMEM so the tensor pipe runs ahead of the scan.
//  * Each epilogue thread owns one user row x NT columns and keeps that row's running top-10 in registers:
//    group maxima (FMNMX3) are compared with the current 10th best first; candidate columns are found with a
//    warp-wide OR of compare masks and re-read from TMEM one column at a time, so insertion code runs rarely.
// n_split = 2/3 feeds hi/lo bf16 splits of the fp32 user features as extra K (near-fp32 scores).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

static constexpr int TK = 10;               // list length kept per row
static constexpr int TILE_U = 128;          // users per MMA (TMEM lanes)
static constexpr int KB = 64;               // k-block (bf16 elements)

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
  int debug;              // SRFRD_TOPK_DEBUG (profiling experiments only): 1 = no scan, 2 = no TMEM loads either
  const bf16* feats; int ld_feats;
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

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int UBS, int NT>
__global__ void __launch_bounds__(96 + 128 * UBS, 1)
catalogue_topk_kernel(const __grid_constant__ CUtensorMap tmE, TopkShape s) {
  constexpr int NG = NT / 32;                         // 32-column groups per thread per tile
  constexpr int ACC_STRIDE = UBS * NT;                // TMEM columns per accumulator stage
  constexpr int B_TILE_BYTES = NT * KB * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smB = smem;
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
  const int a_cols = s.n_split * (s.D / 2);           // 32-bit TMEM columns of one user tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmE);
    for (int i = 0; i < s.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * UBS); }
    mbar_init(afull, 4 * UBS);
    mbar_init(aempty, 2);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tA = tmem_base + 2 * ACC_STRIDE;     // user tiles live after the two accumulator stages

  // Warps 0 and 1 run their loops CONVERGED and only the issuing instructions sit under elect.sync: barrier
  // addresses, descriptors and TMEM addresses then stay in uniform registers (inside an `if (lane == 0)`
  // region the compiler wraps every operand in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop, which made each
  // tcgen05.mma cost > 100 issue cycles).
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int ch = unit % s.chunks;
      const int t0 = ch * s.tiles_per_chunk, t1 = min(s.tiles_total, t0 + s.tiles_per_chunk);
      for (int t = t0; t < t1; ++t) {
        for (int kb = 0; kb < s.kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&full[stage], B_TILE_BYTES);
            tma_load_2d(smB + stage * B_TILE_BYTES, &tmE, &full[stage], kb * KB, s.row_lo + t * NT, SRFRD_EVICT_NORMAL);
          }
          __syncwarp();
          if (++stage == s.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // TWO issuing warps, one per accumulator stage: with K = 64 a tile is only a few hundred MMA cycles, about
    // as long as one warp needs for its barrier waits, fences and commits (measured ~350 cycles per tile), so
    // a single issuer leaves the tensor pipe idle half of the time.  Warp w issues tiles n with n % 2 == w.
    const int w = warp - 1;
    const uint32_t idesc = umma_idesc_bf16(TILE_U, NT, 0, 0);
    const uint64_t bdesc0 = umma_smem_desc(smem_u32(smB), 0, 1024);
    const bool fast = s.D == 64 && s.n_split == 1;
    const uint32_t tacc = tmem_base + w * ACC_STRIDE;
    int stage = 0; uint32_t phase = 0, uphase = 0, aphase = 0;
    int n = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int ch = unit % s.chunks;
      const int t0 = ch * s.tiles_per_chunk, t1 = min(s.tiles_total, t0 + s.tiles_per_chunk);
      mbar_wait(afull, uphase);                         // user tiles of this unit are in TMEM
      uphase ^= 1;
      tc_fence_after();
      for (int t = t0; t < t1; ++t, ++n) {
        if ((n & 1) != w) {                             // the other warp's tile: just advance the smem ring
          stage += s.kblocks;
          if (stage >= s.stages) { stage -= s.stages; phase ^= 1; }
          continue;
        }
        if (s.debug != 3) mbar_wait(&tempty[w], aphase ^ 1);
        aphase ^= 1;
        tc_fence_after();
        for (int kb = 0; kb < s.kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t bd = bdesc0 + (uint64_t)(stage * (B_TILE_BYTES >> 4));
          if (elect_one()) {
            if (fast) {                                   // D == 64, one split: 4 K-steps, everything unrolled
#pragma unroll
              for (int ub = 0; ub < UBS; ++ub)           // the item tile is reused for every user tile
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_ts(tacc + ub * NT, tA + ub * 32 + k * 8, bd + 2 * k, idesc, k != 0);
            } else {
              const int ksteps = min(KB / 16, (s.D - kb * KB + 15) / 16);
              for (int ub = 0; ub < UBS; ++ub)
                for (int sp = 0; sp < s.n_split; ++sp)
                  for (int k = 0; k < ksteps; ++k)
                    umma_bf16_ts(tacc + ub * NT, tA + ub * a_cols + sp * (s.D / 2) + (kb * (KB / 16) + k) * 8,
                                 bd + 2 * k, idesc, (kb | sp | k) != 0);
            }
            umma_commit(&empty[stage]);
            if (kb == s.kblocks - 1) umma_commit(&tfull[w]);
          }
          __syncwarp();
          if (++stage == s.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (elect_one()) umma_commit(aempty);             // this warp's MMAs of the unit are done with the user tiles
      __syncwarp();
    }
  } else {
    const int e = warp - 3;
    const int quarter = warp & 3, ub = e >> 2;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    int as = 0; uint32_t aphase = 0, uphase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int ug = unit / s.chunks, ch = unit % s.chunks;
      const int t0 = ch * s.tiles_per_chunk, t1 = min(s.tiles_total, t0 + s.tiles_per_chunk);
      const int urow = (ug * UBS + ub) * TILE_U + quarter * 32 + lane;
      // ---- stage this thread's user row into tensor memory (A operand, K-major, 2 bf16 per column) ----
      mbar_wait(aempty, uphase ^ 1);                      // previous unit's MMAs no longer read the user tiles
      uphase ^= 1;
      tc_fence_after();
      for (int sp = 0; sp < s.n_split; ++sp) {
        const uint4* src = reinterpret_cast<const uint4*>(s.feats + ((size_t)sp * s.u_pad + urow) * s.ld_feats);
        for (int c = 0; c < s.D / 2; c += 8) {            // 8 columns = 16 features = one UMMA K step
          uint32_t v[8];
          uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
          if (urow < s.U) { lo = __ldg(src + c / 4); hi = __ldg(src + c / 4 + 1); }
          v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
          tmem_st8(tA + lane_off + ub * a_cols + sp * (s.D / 2) + c, v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(afull);

      float ts[TK]; int ti[TK];
#pragma unroll
      for (int r = 0; r < TK; ++r) { ts[r] = -INFINITY; ti[r] = -1; }
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&tfull[as], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_off + as * ACC_STRIDE + ub * NT;
        const int col_row0 = s.row_lo + t * NT;                      // local table row of column 0
        // all NT scores of this thread's row in flight at once (one TMEM round trip per tile)
        uint32_t raw[NG][32];
        if (s.debug == 2) {
#pragma unroll
          for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 32; ++j) raw[g][j] = 0xff800000u;
        } else {
#pragma unroll
          for (int g = 0; g < NG; ++g) tmem_ld32(taddr + g * 32, raw[g]);
          tmem_ld_wait();
        }
        if (s.debug >= 1) {
          uint32_t x = 0;
#pragma unroll
          for (int g = 0; g < NG; ++g) x |= raw[g][0] & raw[g][31];
          if (x == 0x12345678u) ts[0] = 1.f;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[as]);
          if (++as == 2) { as = 0; aphase ^= 1; }
          continue;
        }
        float gm[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
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
        float tmax = gm[0];
#pragma unroll
        for (int g = 1; g < NG; ++g) tmax = fmaxf(tmax, gm[g]);
        if (__any_sync(0xffffffffu, tmax > ts[TK - 1])) {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
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

// tile configuration: UBS user tiles x NT item columns; TMEM: 2 * UBS * NT accumulator + UBS * a_cols operand columns
struct TopkCfg { int ubs, nt; };
static TopkCfg pick_cfg(int D, int n_split) {
  const int a_cols = n_split * (D / 2);
  const TopkCfg cands[] = {{2, 96}, {3, 64}, {2, 64}, {1, 128}, {1, 64}, {1, 32}};   // measured order at C3
  const char* force = getenv("SRFRD_TOPK_CFG");       // e.g. "2x96" (experiments)
  if (force) {
    int u = 0, n = 0;
    if (sscanf(force, "%dx%d", &u, &n) == 2)
      for (const TopkCfg& c : cands)
        if (c.ubs == u && c.nt == n && 2 * c.ubs * c.nt + c.ubs * a_cols <= 512) return c;
  }
  for (const TopkCfg& c : cands)
    if (2 * c.ubs * c.nt + c.ubs * a_cols <= 512) return c;
  return TopkCfg{0, 0};
}

static int plan_chunks(const TopkCfg& c, int64_t U, int64_t n_rows, int64_t row_lo) {
  const int64_t tiles = (n_rows - row_lo + c.nt - 1) / c.nt;
  const int64_t ugroups = (U + c.ubs * TILE_U - 1) / (c.ubs * TILE_U);
  // One wave of work units: every unit restarts its top-10 lists cold (about 10 ln(n/10) insertions per row
  // over n items), so item chunks are only used to occupy SMs that the user groups alone would leave idle.
  int64_t chunks = ugroups > 0 ? num_sms() / ugroups : 1;
  if (chunks > tiles / 8) chunks = tiles / 8;
  if (chunks < 1) chunks = 1;
  const int64_t per = (tiles + chunks - 1) / chunks;
  chunks = per > 0 ? (tiles + per - 1) / per : 1;
  return (int)(chunks < 1 ? 1 : chunks);
}

extern "C" int srfrd_catalogue_topk_plan(int64_t U, int64_t n_rows, int64_t row_lo, int D, int n_split, int* chunks_out) {
  SRFRD_REQUIRE(chunks_out, "catalogue_topk_plan: null output");
  const TopkCfg c = pick_cfg(D, n_split);
  SRFRD_REQUIRE(c.ubs > 0, "catalogue_topk: D=%d with n_split=%d does not fit tensor memory", D, n_split);
  *chunks_out = plan_chunks(c, U, n_rows, row_lo);
  return 0;
}

template <int UBS, int NT>
static int launch_topk(const CUtensorMap& tmE, TopkShape& s, cudaStream_t stream) {
  const int b_tile = NT * KB * 2;
  s.stages = (200 * 1024) / b_tile;
  if (s.stages > 12) s.stages = 12;
  const size_t smem = (size_t)s.stages * b_tile + 1024 + 512;
  static bool attr_set = false;
  if (!attr_set) {
    SRFRD_CUDA(cudaFuncSetAttribute(catalogue_topk_kernel<UBS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int grid = s.ugroups * s.chunks;
  if (grid > num_sms()) grid = num_sms();
  catalogue_topk_kernel<UBS, NT><<<grid, 96 + 128 * UBS, smem, stream>>>(tmE, s);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_catalogue_topk(const void* feats_bf16, int64_t U, int64_t u_pad, int n_split, const void* table_bf16,
                                    int64_t n_rows, int64_t row_lo, int64_t id_base, int D, int ld_feats, int ld_table,
                                    int chunks, float* part_scores, int* part_ids, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRFRD_REQUIRE(feats_bf16 && table_bf16 && part_scores && part_ids, "catalogue_topk: null pointer");
  SRFRD_REQUIRE(n_split >= 1 && n_split <= 3, "catalogue_topk: n_split must be 1..3");
  SRFRD_REQUIRE(D % 16 == 0 && ld_feats % 8 == 0 && ld_table % 8 == 0, "catalogue_topk: D %% 16 and ld %% 8 required (D=%d)", D);
  SRFRD_REQUIRE(((uintptr_t)feats_bf16 & 15) == 0, "catalogue_topk: feats must be 16-byte aligned");
  SRFRD_REQUIRE(U > 0 && n_rows > row_lo && row_lo >= 0, "catalogue_topk: empty problem");
  SRFRD_REQUIRE(n_rows < (1ll << 31) && id_base + n_rows < (1ll << 31), "catalogue_topk: ids must fit int32");
  const TopkCfg c = pick_cfg(D, n_split);
  SRFRD_REQUIRE(c.ubs > 0, "catalogue_topk: D=%d with n_split=%d does not fit tensor memory", D, n_split);
  SRFRD_REQUIRE(chunks == plan_chunks(c, U, n_rows, row_lo), "catalogue_topk: chunks must come from srfrd_catalogue_topk_plan");
  SRFRD_REQUIRE(u_pad >= U, "catalogue_topk: u_pad < U");
  TopkShape s;
  s.U = (int)U; s.D = D; s.n_split = n_split; s.u_pad = (int)u_pad;
  s.row_lo = (int)row_lo; s.row_hi = (int)n_rows; s.id_base = id_base;
  s.kblocks = (D + KB - 1) / KB;
  s.tiles_total = (int)((n_rows - row_lo + c.nt - 1) / c.nt);
  s.chunks = chunks;
  s.tiles_per_chunk = (s.tiles_total + chunks - 1) / chunks;
  s.ugroups = (int)((U + c.ubs * TILE_U - 1) / (c.ubs * TILE_U));
  s.feats = (const bf16*)feats_bf16; s.ld_feats = ld_feats;
  s.out_scores = part_scores; s.out_ids = part_ids;
  { const char* dbg = getenv("SRFRD_TOPK_DEBUG"); s.debug = dbg ? atoi(dbg) : 0; }
  CUtensorMap tmE;
  if (int rc = make_tmap_bf16_2d(&tmE, table_bf16, n_rows, D, ld_table, c.nt, KB)) return rc;
  if (c.ubs == 3 && c.nt == 64) return launch_topk<3, 64>(tmE, s, stream);
  if (c.ubs == 2 && c.nt == 96) return launch_topk<2, 96>(tmE, s, stream);
  if (c.ubs == 2 && c.nt == 64) return launch_topk<2, 64>(tmE, s, stream);
  if (c.ubs == 1 && c.nt == 128) return launch_topk<1, 128>(tmE, s, stream);
  if (c.ubs == 1 && c.nt == 64) return launch_topk<1, 64>(tmE, s, stream);
  return launch_topk<1, 32>(tmE, s, stream);
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

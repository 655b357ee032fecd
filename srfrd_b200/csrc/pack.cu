// Packed token layout: stop computing padding (SRFR_model.py:98-99, :113, :121).
//
// In the reference every (b, l) slot is a token, but the sampler's batches are ~88 % padding (C2): a pad slot enters the
// encoder as x = 0 (pad mask, SRFR_model.py:98-99), is re-zeroed after every block (:121), and there is NO key-padding
// mask (:113 is commented out), so it still acts as a KEY / VALUE with k = b_k, v = b_v -- identical for every pad slot
// of a sequence.  As a QUERY a pad slot is dead inside the blocks (its output is masked), and after the final LayerNorm
// its hidden state is one constant row.  So the packed layout keeps, per sequence,
//     [ one PAD REPRESENTATIVE row ] [ the kept tokens in position order ]
// where a token is kept if its input id != 0 (or its `keep` id != 0: positions that carry a loss term even though their
// input is a pad).  The representative stands for all DROPPED pad slots: row-wise kernels (LayerNorm, GEMMs) treat it as
// an ordinary row with id 0; attention gives the key column of the representative the weight
//     mult(i) = number of dropped pad slots at positions < position(i)
// in query i's softmax numerator and denominator (p copies of exp(s) == p * exp(s)), which is exact, also for the
// gradients (the representative's dk / dv rows are the sums over the copies, and its x = 0 adds nothing to dW).
// Filler rows round the row count up to a multiple of 128 (one GEMM / weight-gradient tile): single-row sequences with
// id 0, so every kernel finds finite, well-defined values in every row it touches and their gradients are exactly 0.
// Everything is computed on the device (the row count is data dependent; the step runs inside a CUDA graph):
//   pack_count : kept tokens per sequence (one warp per sequence)
//   pack_scan  : ONE block -- exclusive scan -> first row of every sequence, total rows M, filler rows, and the greedy
//                attention tile plan (tiles of <= 128 rows that hold whole sequences; found by parallel binary searches
//                for every possible tile start followed by one short pointer chase through shared memory)
//   pack_fill  : per-row maps (dense token, item id, sequence bounds, pad multiplicity, position) and the inverse map
#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

static constexpr int PACK_TILE = 128;
static constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(256) pack_count_kernel(const int64_t* seq, const int64_t* keep, int64_t B, int L, int* cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int n = 0;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    bool k = false;
    if (l < L) k = (__ldg(seq + b * L + l) != 0) || (keep && __ldg(keep + b * L + l) != 0);
    n += __popc(__ballot_sync(0xffffffffu, k));
  }
  if (lane == 0) cnt[b] = n;
}

// bound(i): first row of "sequence" i, where the filler rows count as single-row sequences after the B real ones
__device__ __forceinline__ int pack_bound(const int* first, int B, int Tp, int i) { return i <= B ? first[min(i, B)] + 0 : Tp + (i - B); }

__global__ void __launch_bounds__(SCAN_THREADS) pack_scan_kernel(srfrd_pack_t pk, int64_t B, int plan_tiles) {
  extern __shared__ int sm[];                 // [B + 1] first rows, then [NBmax] next-tile pointers
  __shared__ int wsum[32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* first = sm;
  // ---- exclusive scan of (cnt[b] + 1): each thread owns a run of consecutive sequences
  const int per = (int)((B + SCAN_THREADS - 1) / SCAN_THREADS);
  const int64_t b0 = (int64_t)tid * per, b1 = min(B, b0 + per);
  int local = 0;
  for (int64_t b = b0; b < b1; ++b) local += pk.cnt[b] + 1;
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
    if (lane == 31) s_total = v;
  }
  __syncthreads();
  int run = incl - local + (warp ? wsum[warp - 1] : 0);
  for (int64_t b = b0; b < b1; ++b) { first[b] = run; pk.seq_first[b] = run; run += pk.cnt[b] + 1; }
  const int Tp = s_total;
  const int M = (Tp + PACK_TILE - 1) / PACK_TILE * PACK_TILE;
  if (tid == 0) { first[B] = Tp; pk.seq_first[B] = Tp; pk.rows[0] = M; pk.rows[1] = Tp; }
  __syncthreads();
  // ---- filler rows [T', M): single-row sequences with id 0
  for (int r = Tp + tid; r < M; r += SCAN_THREADS) {
    pk.row_tok[r] = -1;
    pk.row_ids[r] = 0;
    reinterpret_cast<int4*>(pk.row_info)[r] = make_int4(r, r + 1, __float_as_int(1.f), 0);
  }
  // ---- greedy tile plan.  A tile starts at a sequence boundary and holds every following whole sequence that fits in
  // 128 rows.  nxt[i] = index of the boundary where the tile that starts at boundary i ends (largest k with
  // bound(k) <= bound(i) + 128): one binary search per boundary, all in parallel; the plan is then a pointer chase.
  if (!plan_tiles) {                          // sequences longer than a tile: row maps only (dense-layout attention)
    if (tid == 0) { pk.rows[2] = 0; pk.rows[3] = 0; pk.tile_row0[0] = M; }
    return;
  }
  const int NB = (int)B + (M - Tp);           // boundaries 0 .. NB (bound(NB) = M)
  int* nxt = sm + B + 1;
  for (int i = tid; i < NB; i += SCAN_THREADS) {
    const int lim = pack_bound(first, (int)B, Tp, i) + PACK_TILE;
    int lo = i + 1, hi = NB;                  // bound(i + 1) <= lim always (a sequence has at most 128 rows)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (pack_bound(first, (int)B, Tp, mid) <= lim) lo = mid; else hi = mid - 1;
    }
    nxt[i] = lo;
  }
  __syncthreads();
  // the chase itself touches shared memory only (one dependent load per tile); the visited boundaries are recorded in
  // place (the slots of nxt[] already passed are dead) and written out by all threads afterwards
  __shared__ int s_ntiles;
  if (tid == 0) {
    int k = 0;
    for (int i = 0; i < NB;) { const int j = nxt[i]; nxt[k++] = i; i = j; }      // k <= i always: slot k is already consumed
    s_ntiles = k;
    pk.rows[2] = k;
    pk.rows[3] = 0;
  }
  __syncthreads();
  const int nt = s_ntiles;
  for (int k = tid; k < nt; k += SCAN_THREADS) pk.tile_row0[k] = pack_bound(first, (int)B, Tp, nxt[k]);
  if (tid == 0) pk.tile_row0[nt] = M;
}

__global__ void __launch_bounds__(256) pack_fill_kernel(const int64_t* seq, const int64_t* keep, int64_t B, int L, srfrd_pack_t pk) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int row0 = pk.seq_first[b], end = pk.seq_first[b + 1];
  int4* info = reinterpret_cast<int4*>(pk.row_info);
  if (lane == 0) {
    pk.row_tok[row0] = -1;
    pk.row_ids[row0] = 0;
    info[row0] = make_int4(row0, end, __float_as_int(1.f), 0);
  }
  int base = 0;                               // kept tokens at positions < l0
  int last = row0;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    int64_t id = 0;
    bool k = false;
    if (l < L) {
      id = __ldg(seq + b * L + l);
      k = (id != 0) || (keep && __ldg(keep + b * L + l) != 0);
    }
    const unsigned m = __ballot_sync(0xffffffffu, k);
    const int before = base + __popc(m & ((1u << lane) - 1u));       // kept tokens at positions < l
    const int row = row0 + 1 + before;
    if (l < L) {
      pk.tok_row[b * L + l] = k ? row : -1;
      if (k) {
        pk.row_tok[row] = (int)(b * L + l);
        pk.row_ids[row] = id;
        info[row] = make_int4(row0, end, __float_as_int((float)(l - before)), l);     // l - before = dropped pads before l
        if (l == L - 1) last = row;
      }
    }
    base += __popc(m);
  }
  last = __shfl_sync(0xffffffffu, last, (L - 1) & 31);
  if (lane == 0) pk.last_row[b] = last;
}

}  // namespace srfrd

using namespace srfrd;

extern "C" int srfrd_pack_plan(const int64_t* seq, const int64_t* keep, int64_t B, int L, const srfrd_pack_t* pk_host,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRFRD_REQUIRE(seq && pk_host, "pack_plan: null pointer");
  const srfrd_pack_t pk = *pk_host;
  SRFRD_REQUIRE(pk.rows && pk.cnt && pk.seq_first && pk.tok_row && pk.row_tok && pk.row_ids && pk.row_info && pk.tile_row0 &&
                pk.last_row, "pack_plan: null output");
  SRFRD_REQUIRE(B >= 1 && L >= 1 && L <= 4096, "pack_plan: bad shape (B=%lld L=%d)", (long long)B, L);
  SRFRD_REQUIRE(B <= 16384, "pack_plan: at most 16384 sequences per call (B=%lld)", (long long)B);
  const int64_t need = (B * (L + 1) + PACK_TILE - 1) / PACK_TILE * PACK_TILE;
  SRFRD_REQUIRE(pk.cap >= need, "pack_plan: capacity %lld rows < %lld", (long long)pk.cap, (long long)need);
  SRFRD_REQUIRE(((uintptr_t)pk.row_info & 15) == 0, "pack_plan: row_info must be 16-byte aligned");
  pack_count_kernel<<<(unsigned)((B + 7) / 8), 256, 0, stream>>>(seq, keep, B, L, pk.cnt);
  SRFRD_LAUNCH_CHECK();
  const size_t smem = (size_t)(2 * B + 2 + PACK_TILE + 2) * sizeof(int);
  static bool attr = false;
  if (!attr) {
    SRFRD_CUDA(cudaFuncSetAttribute(pack_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  pack_scan_kernel<<<1, SCAN_THREADS, smem, stream>>>(pk, B, L + 1 <= PACK_TILE ? 1 : 0);
  SRFRD_LAUNCH_CHECK();
  pack_fill_kernel<<<(unsigned)((B + 7) / 8), 256, 0, stream>>>(seq, keep, B, L, pk);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

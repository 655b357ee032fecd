// Moving rows between the packed token layout (pack.cu) and the dense (B, L) layout, for sequence lengths whose attention
// kernels work on the dense layout (128 < maxlen <= 256: attention_long.cu).  The row-wise bulk of an encoder block
// (projections, FFN, LayerNorms: ~2/3 of the C4 step) then runs on the packed rows only, and attention sees the dense
// tensors it expects:
//   unpack : dense[t] = packed[tok_row[t]];  a dropped pad slot takes its sequence's pad-representative row (q, k, v: the
//            values every pad slot has -- k = b_k, v = b_v) or zeros (dO: a pad query's output is dead)
//   pack   : packed[r] = dense[row_tok[r]];  a pad-representative row takes zeros (o, dq) or the SUM over its sequence's
//            dropped pad slots (dk, dv: every copy of the pad key contributes), filler rows zeros
// 16-byte vectors, one warp per row (coalesced), HBM-bound copies.
#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

struct RepackPart {
  const bf16* src; int ld_s;
  bf16* dst; int ld_d;
  int W;                 // columns (multiple of 8)
  int mode;              // unpack: 0 = pad slots <- representative row, 1 = pad slots <- 0
                         // pack:   0 = representative rows <- 0,           1 = representative rows <- sum of dropped pad slots
};
struct RepackParams {
  RepackPart part[3];
  int n_parts;
  const int* tok_row;    // (B * L) packed row of each dense token, -1 = dropped pad
  const int* row_tok;    // (cap) dense token of each packed row, -1 = representative / filler
  const int* seq_first;  // (B + 1)
  const int* rows_dev;   // {M, T', ...}
  int64_t B; int L; int64_t cap;
};

__global__ void __launch_bounds__(256) unpack_rows_kernel(RepackParams p) {
  pdl_prologue_done();
  const int lane = threadIdx.x & 31;
  const int64_t T = p.B * p.L;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < T; t += nw) {
    const int r = __ldg(p.tok_row + t);
    const int rep = __ldg(p.seq_first + t / p.L);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k >= p.n_parts) break;
      const RepackPart& q = p.part[k];
      const bool zero = r < 0 && q.mode == 1;
      const uint4* s = reinterpret_cast<const uint4*>(q.src + (int64_t)(r >= 0 ? r : rep) * q.ld_s);
      uint4* d = reinterpret_cast<uint4*>(q.dst + t * q.ld_d);
      for (int c = lane; c < q.W / 8; c += 32) d[c] = zero ? make_uint4(0, 0, 0, 0) : __ldg(s + c);
    }
  }
}

__global__ void __launch_bounds__(256) pack_rows_kernel(RepackParams p) {
  pdl_prologue_done();
  const int lane = threadIdx.x & 31;
  const int64_t M = min(p.cap, (int64_t)__ldg(p.rows_dev));
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < M; r += nw) {
    const int tok = __ldg(p.row_tok + r);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k >= p.n_parts) break;
      const RepackPart& q = p.part[k];
      uint4* d = reinterpret_cast<uint4*>(q.dst + r * q.ld_d);
      if (tok >= 0) {
        const uint4* s = reinterpret_cast<const uint4*>(q.src + (int64_t)tok * q.ld_s);
        for (int c = lane; c < q.W / 8; c += 32) d[c] = __ldg(s + c);
      } else if (q.mode == 0) {
        for (int c = lane; c < q.W / 8; c += 32) d[c] = make_uint4(0, 0, 0, 0);
      }
      // (mode 1: representative rows are written by pack_pad_sum_kernel, filler rows below)
    }
  }
}

// representative row of sequence b <- sum over the dropped pad slots of b (fp32 accumulation); filler rows <- 0.
// One block per sequence: the dropped positions are compacted into shared memory first, then every warp sums a stripe of
// them for all columns (lane = 16-byte chunk), and the warps' partial sums meet in shared memory.
static constexpr int PADSUM_WARPS = 8;
static constexpr int PADSUM_MAXCH = 3;            // 16-byte chunks per lane: rows up to 3 * 32 * 8 = 768 columns
__global__ void __launch_bounds__(32 * PADSUM_WARPS) pack_pad_sum_kernel(RepackParams p) {
  extern __shared__ float sacc[];                // [PADSUM_WARPS][W] partial sums, then int pad list [L]
  __shared__ int s_n;
  pdl_prologue_done();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t M = min(p.cap, (int64_t)__ldg(p.rows_dev));
  const int Tp = __ldg(p.rows_dev + 1);
  int Wmax = 0;
  for (int k = 0; k < p.n_parts; ++k) if (p.part[k].mode == 1) Wmax = max(Wmax, p.part[k].W);
  int* pads = reinterpret_cast<int*>(sacc + PADSUM_WARPS * Wmax);
  for (int64_t b = blockIdx.x; b < p.B + (M - Tp); b += gridDim.x) {
    if (b >= p.B) {                                      // filler row
      for (int k = 0; k < p.n_parts; ++k) {
        const RepackPart& q = p.part[k];
        if (q.mode != 1) continue;
        uint4* d = reinterpret_cast<uint4*>(q.dst + (Tp + (b - p.B)) * q.ld_d);
        for (int c = threadIdx.x; c < q.W / 8; c += blockDim.x) d[c] = make_uint4(0, 0, 0, 0);
      }
      continue;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int l = threadIdx.x; l < p.L; l += blockDim.x)
      if (__ldg(p.tok_row + b * p.L + l) < 0) pads[atomicAdd(&s_n, 1)] = l;       // (order is irrelevant for a sum... of
    __syncthreads();                                                             //  fp32 partials rounded once at the end)
    const int n = s_n;
    const int rep = __ldg(p.seq_first + b);
    for (int k = 0; k < p.n_parts; ++k) {
      const RepackPart& q = p.part[k];
      if (q.mode != 1) continue;
      const int nch = q.W / 8;
      float acc[PADSUM_MAXCH][8];
#pragma unroll
      for (int u = 0; u < PADSUM_MAXCH; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[u][j] = 0.f;
      for (int i = warp; i < n; i += PADSUM_WARPS) {
        const uint4* srow = reinterpret_cast<const uint4*>(q.src + (b * p.L + pads[i]) * q.ld_s);
#pragma unroll
        for (int u = 0; u < PADSUM_MAXCH; ++u) {
          const int c = u * 32 + lane;
          if (c < nch) {
            const uint4 v = __ldg(srow + c);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
              acc[u][2 * j] += f.x; acc[u][2 * j + 1] += f.y;
            }
          }
        }
      }
      __syncthreads();                                   // (previous part's reduction has been read)
#pragma unroll
      for (int u = 0; u < PADSUM_MAXCH; ++u) {
        const int c = u * 32 + lane;
        if (c < nch)
#pragma unroll
          for (int j = 0; j < 8; ++j) sacc[warp * q.W + c * 8 + j] = acc[u][j];
      }
      __syncthreads();
      for (int c = threadIdx.x; c < nch; c += blockDim.x) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < PADSUM_WARPS; ++w2) a += sacc[w2 * q.W + c * 8 + j];
          t[j] = a;
        }
        reinterpret_cast<uint4*>(q.dst + (int64_t)rep * q.ld_d)[c] =
            make_uint4(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]), pack_bf16x2(t[4], t[5]), pack_bf16x2(t[6], t[7]));
      }
    }
  }
}

}  // namespace srfrd

using namespace srfrd;

static int fill_parts(RepackParams& p, const srfrd_repack_part_t* parts, int n_parts, const char* who) {
  SRFRD_REQUIRE(parts && n_parts >= 1 && n_parts <= 3, "%s: 1..3 parts", who);
  p.n_parts = n_parts;
  for (int k = 0; k < n_parts; ++k) {
    SRFRD_REQUIRE(parts[k].src && parts[k].dst && parts[k].W % 8 == 0 && parts[k].ld_s % 8 == 0 && parts[k].ld_d % 8 == 0 &&
                      parts[k].ld_s >= parts[k].W && parts[k].ld_d >= parts[k].W,
                  "%s: part %d needs 16-byte aligned rows (W=%d)", who, k, parts[k].W);
    SRFRD_REQUIRE((((uintptr_t)parts[k].src | (uintptr_t)parts[k].dst) & 15) == 0, "%s: part %d is not 16-byte aligned", who, k);
    p.part[k].src = (const bf16*)parts[k].src; p.part[k].ld_s = parts[k].ld_s; p.part[k].dst = (bf16*)parts[k].dst;
    p.part[k].ld_d = parts[k].ld_d; p.part[k].W = parts[k].W; p.part[k].mode = parts[k].mode;
  }
  return 0;
}

extern "C" int srfrd_unpack_rows(const srfrd_repack_part_t* parts, int n_parts, const srfrd_pack_t* pk, int64_t B, int L,
                                 void* stream) {
  SRFRD_REQUIRE(pk && pk->tok_row && pk->seq_first, "unpack_rows: null plan");
  RepackParams p = {};
  if (int rc = fill_parts(p, parts, n_parts, "unpack_rows")) return rc;
  p.tok_row = pk->tok_row; p.row_tok = pk->row_tok; p.seq_first = pk->seq_first; p.rows_dev = pk->rows; p.B = B; p.L = L; p.cap = pk->cap;
  if (B * L == 0) return 0;
  int64_t grid = (B * L + 7) / 8;
  if (grid > (int64_t)num_sms() * 16) grid = (int64_t)num_sms() * 16;
  SRFRD_CUDA(launch_pdl(unpack_rows_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, p));
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_pack_rows(const srfrd_repack_part_t* parts, int n_parts, const srfrd_pack_t* pk, int64_t B, int L,
                               void* stream) {
  SRFRD_REQUIRE(pk && pk->tok_row && pk->row_tok && pk->seq_first && pk->rows, "pack_rows: null plan");
  RepackParams p = {};
  if (int rc = fill_parts(p, parts, n_parts, "pack_rows")) return rc;
  p.tok_row = pk->tok_row; p.row_tok = pk->row_tok; p.seq_first = pk->seq_first; p.rows_dev = pk->rows; p.B = B; p.L = L; p.cap = pk->cap;
  if (B * L == 0) return 0;
  int64_t grid = (pk->cap + 7) / 8;
  if (grid > (int64_t)num_sms() * 16) grid = (int64_t)num_sms() * 16;
  SRFRD_CUDA(launch_pdl(pack_rows_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, p));
  SRFRD_LAUNCH_CHECK();
  bool any_sum = false;
  for (int k = 0; k < n_parts; ++k) any_sum |= parts[k].mode == 1;
  if (any_sum) {
    int Wmax = 0;
    for (int k = 0; k < n_parts; ++k) if (parts[k].mode == 1) Wmax = parts[k].W > Wmax ? parts[k].W : Wmax;
    SRFRD_REQUIRE(Wmax <= PADSUM_MAXCH * 256, "pack_rows: rows wider than %d columns are not supported", PADSUM_MAXCH * 256);
    const size_t smem = (size_t)PADSUM_WARPS * Wmax * sizeof(float) + (size_t)L * sizeof(int);
    SRFRD_REQUIRE(smem <= 48 * 1024, "pack_rows: W = %d, L = %d exceed the shared-memory budget of the pad sum", Wmax, L);
    int64_t g2 = B + 127;
    if (g2 > (int64_t)num_sms() * 8) g2 = (int64_t)num_sms() * 8;
    SRFRD_CUDA(launch_pdl(pack_pad_sum_kernel, dim3((unsigned)g2), dim3(32 * PADSUM_WARPS), smem, (cudaStream_t)stream, p));
    SRFRD_LAUNCH_CHECK();
  }
  return 0;
}

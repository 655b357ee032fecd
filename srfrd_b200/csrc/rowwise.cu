// HBM-bound row-wise kernels: K1 (gather + positional add + concat/add + pad mask + LayerNorm),
// LayerNorm forward/backward, column sums (bias gradients), fp32 -> bf16 weight shadows.
//
// Layout: a row (token) is handled by a group of LPR lanes (8/16/32), each lane owning CH chunks of
// 8 contiguous features, so every global access is a 16-byte (bf16) or 2 x 16-byte (fp32) vector and a
// warp covers 32/LPR rows per pass; two passes are kept in flight per loop iteration for memory-level
// parallelism.  Row statistics are reduced with xor-shuffles inside the lane group.
#include <stdlib.h>

#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

static constexpr int MAXW = 512;     // widest supported row

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void load8_f32(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// item-table rows are read once per lookup: no L1 allocation, 256-byte L2 fetch granularity (a D = 64 row is 256 B)
__device__ __forceinline__ void load8_f32_stream(const float* p, float* f) {
  float4 a, b;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8_f32(float* p, const float* f) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// ---------------------------------------------------------------------------------------------
struct EmbedParams {
  const float* item_table;   // (n_rows, D)
  const float* pos_table;    // (L, D)
  const float* aux_table;    // mode 1: fake_embed (3, F); mode 2: user_label_embed (labels, D)
  const int64_t* seq;        // (B, L)
  const int64_t* aux_ids;    // mode 1: (B, L) fake ids or null (-> all 0); mode 2: (B,) labels
  const int* row_tok;        // packed layout: dense token (b * L + l) of output row t, -1 = pad representative / filler
  const int* rows_dev;       // packed layout: device row count (rows beyond it are not touched)
  int64_t T;
  int L, D, F, mode;
  float item_scale;
  const float* ln_w;         // (H) or null -> no LN output
  const float* ln_b;
  float eps;
  bf16* x0;                  // (T, ldx) or null
  float* x0_f32;             // (T, H) or null
  bf16* q;                   // (T, ldx) LN output or null
  float* stats;              // (T, 2) mean, rstd or null
  int ldx;
  uint64_t drop_seed; uint32_t drop_thresh, drop_stream; float drop_scale; const float* drop_step;
};

// VEC: D (and F) multiples of 8 -> 16-byte vector loads; otherwise (the reference's own 45+5 / 50+10 widths) the same
// arithmetic element by element.  Columns >= H (up to the padded leading dimension) are written as zeros.
template <int LPR, int CH, bool VEC>
__global__ void __launch_bounds__(256, 4) embed_ln_kernel(EmbedParams p) {
  pdl_prologue_done();
  if (p.drop_thresh) p.drop_seed = mix_seed(p.drop_seed, p.drop_step);
  constexpr int RPW = 32 / LPR;                         // rows per warp pass
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  const int H = p.D + (p.mode == 1 ? p.F : 0);
  const int Hc = (H + 7) & ~7;                          // chunked width (8 columns per lane chunk)
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t0 = warp * RPW; t0 < p.T; t0 += nwarps * RPW) {
    const int64_t t = t0 + grp;
    const bool row_ok = t < p.T;
    int64_t id = 0, aid = 0;
    int l = 0;
    if (row_ok) {
      l = (int)(t % p.L);
      id = __ldg(p.seq + t);
      if (p.mode == 1) aid = p.aux_ids ? __ldg(p.aux_ids + t) : 0;
      if (p.mode == 2) aid = __ldg(p.aux_ids + t / p.L);
    }
    const bool valid = row_ok && id != 0;
    float v[CH][8];
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c = (ch * LPR + sub) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[ch][j] = 0.f;
      if (!VEC) {
        if (valid && c < Hc) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = c + j;
            float x = 0.f;
            if (col < p.D) {
              x = __ldg(p.item_table + id * p.D + col);
              if (p.item_scale != 1.f) x = __fmul_rn(x, p.item_scale);
              x = __fadd_rn(x, __ldg(p.pos_table + (int64_t)l * p.D + col));
              if (p.mode == 2) x = __fadd_rn(x, __ldg(p.aux_table + aid * p.D + col));
            } else if (col < H) {
              x = __ldg(p.aux_table + aid * p.F + (col - p.D));
            }
            if (p.drop_thresh && col < H)
              x = dropout_keep(p.drop_seed, p.drop_stream, (uint64_t)t * H + col, p.drop_thresh) ? x * p.drop_scale : 0.f;
            v[ch][j] = x;
          }
        }
      } else if (valid && c < H) {
        if (c < p.D) {
          // __fmul_rn/__fadd_rn: no FMA contraction, so the pre-LN tensor is bit-identical to
          // torch's  E[id] (* sqrt(d)) + P[l] (+ Ul[label])   (SRFR_model.py:22-25, :622-624, :419-422)
          float e[8], pp[8];
          load8_f32_stream(p.item_table + id * p.D + c, e);
          load8_f32(p.pos_table + (int64_t)l * p.D + c, pp);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x = e[j];
            if (p.item_scale != 1.f) x = __fmul_rn(x, p.item_scale);
            v[ch][j] = __fadd_rn(x, pp[j]);
          }
          if (p.mode == 2) {
            float u[8];
            load8_f32(p.aux_table + aid * p.D + c, u);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[ch][j] = __fadd_rn(v[ch][j], u[j]);
          }
        } else {
          load8_f32(p.aux_table + aid * p.F + (c - p.D), v[ch]);
        }
        if (p.drop_thresh) {   // SASRec.emb_dropout (SRFR_model.py:625): after the positional add, before the mask
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[ch][j] = dropout_keep(p.drop_seed, p.drop_stream, (uint64_t)t * H + c + j, p.drop_thresh)
                           ? v[ch][j] * p.drop_scale : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[ch][j];
    }
    float mean = 0.f, rstd = 0.f;
    if (p.ln_w) {
      mean = group_sum<LPR>(sum) / H;
      float sq = 0.f;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        const int c = (ch * LPR + sub) * 8;
        if (c < Hc) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = v[ch][j] - mean; sq += (VEC || c + j < H) ? d * d : 0.f; }
        }
      }
      rstd = rsqrtf(group_sum<LPR>(sq) / H + p.eps);
    }
    if (!row_ok) continue;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c = (ch * LPR + sub) * 8;
      if (c >= Hc) continue;
      if (p.x0_f32) {
        if (VEC) store8_f32(p.x0_f32 + t * H + c, v[ch]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) if (c + j < H) p.x0_f32[t * H + c + j] = v[ch][j];
        }
      }
      if (p.x0) *reinterpret_cast<uint4*>(p.x0 + t * p.ldx + c) = pack8(v[ch]);
      if (p.ln_w) {
        float w[8], b[8], y[8];
        if (VEC) { load8_f32(p.ln_w + c, w); load8_f32(p.ln_b + c, b); }
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { w[j] = c + j < H ? __ldg(p.ln_w + c + j) : 0.f; b[j] = c + j < H ? __ldg(p.ln_b + c + j) : 0.f; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (VEC || c + j < H) ? (v[ch][j] - mean) * rstd * w[j] + b[j] : 0.f;
        *reinterpret_cast<uint4*>(p.q + t * p.ldx + c) = pack8(y);
      }
    }
    if (p.ln_w && p.stats && sub == 0) *reinterpret_cast<float2*>(p.stats + 2 * t) = make_float2(mean, rstd);
  }
}

// The vector variant (D and F multiples of 8): same arithmetic as above with the embedding mode and the dropout compiled
// in, the position index carried incrementally (no 64-bit division per row), and every row-dependent load issued before
// any arithmetic.  ncu on the generic kernel showed 650 warp instructions per 8 rows (control flow, local-memory spills)
// and the L1 data path at 74 % -- the instruction stream, not HBM, was the bound.
template <int LPR, int CH, int MODE, bool DROP>
__global__ void __launch_bounds__(256, 4) embed_ln_vec_kernel(EmbedParams p) {
  pdl_prologue_done();
  if (DROP) p.drop_seed = mix_seed(p.drop_seed, p.drop_step);
  constexpr int RPW = 32 / LPR;                         // rows per warp pass
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  const int H = p.D + (MODE == 1 ? p.F : 0);
  const float invH = 1.f / (float)H;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5) * RPW;
  const int lstep = (int)(stride % p.L);
  int64_t t = warp * RPW + grp;
  int l = (int)(t % p.L);
  const bool has_ln = p.ln_w != nullptr;
  if (p.rows_dev) p.T = min(p.T, (int64_t)__ldg(p.rows_dev));
  for (; t - grp < p.T; t += stride, l = (l + lstep >= p.L) ? l + lstep - p.L : l + lstep) {
    const bool row_ok = t < p.T;
    int64_t id = 0, aid = 0;
    if (p.row_tok) {                   // packed layout: output row t holds dense token row_tok[t] (or a row with id 0)
      const int src = row_ok ? __ldg(p.row_tok + t) : -1;
      if (src >= 0) {
        l = src % p.L;
        id = __ldg(p.seq + src);
        if (MODE == 1) aid = p.aux_ids ? __ldg(p.aux_ids + src) : 0;
        if (MODE == 2) aid = __ldg(p.aux_ids + src / p.L);
      }
    } else if (row_ok) {
      id = __ldg(p.seq + t);
      if (MODE == 1) aid = p.aux_ids ? __ldg(p.aux_ids + t) : 0;
      if (MODE == 2) aid = __ldg(p.aux_ids + t / p.L);
    }
    const bool valid = row_ok && id != 0;
    const float* erow = p.item_table + id * p.D;
    const float* prow = p.pos_table + (int64_t)l * p.D;
    const float* arow = p.aux_table + aid * (MODE == 1 ? p.F : p.D);
    float v[CH][8];
    float e[CH][8];
    // 1. all gathers in flight
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c = (ch * LPR + sub) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) e[ch][j] = 0.f;
      if (valid && c < H) {
        if (MODE != 1 || c < p.D) load8_f32_stream(erow + c, e[ch]);
        else load8_f32(arow + (c - p.D), e[ch]);
      }
    }
    // 2. positional (and user-label) add: __fmul_rn/__fadd_rn, no FMA contraction, so the pre-LN tensor is bit-identical
    //    to torch's  E[id] (* sqrt(d)) + P[l] (+ Ul[label])   (SRFR_model.py:22-25, :622-624, :419-422)
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c = (ch * LPR + sub) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[ch][j] = e[ch][j];
      if (valid && c < p.D) {
        float pp[8];
        load8_f32(prow + c, pp);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float x = e[ch][j];
          if (p.item_scale != 1.f) x = __fmul_rn(x, p.item_scale);
          v[ch][j] = __fadd_rn(x, pp[j]);
        }
        if (MODE == 2) {
          float u[8];
          load8_f32(arow + c, u);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[ch][j] = __fadd_rn(v[ch][j], u[j]);
        }
      }
      if (DROP) {   // SASRec.emb_dropout (SRFR_model.py:625): after the positional add, before the mask
        if (valid && c < H) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[ch][j] = dropout_keep(p.drop_seed, p.drop_stream, (uint64_t)t * H + c + j, p.drop_thresh)
                           ? v[ch][j] * p.drop_scale : 0.f;
        }
      }
      sum += ((v[ch][0] + v[ch][1]) + (v[ch][2] + v[ch][3])) + ((v[ch][4] + v[ch][5]) + (v[ch][6] + v[ch][7]));
    }
    float a = 0.f, nb = 0.f, mean = 0.f;
    if (has_ln) {
      mean = group_sum<LPR>(sum) * invH;
      float sq = 0.f;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        if ((ch * LPR + sub) * 8 < H) {
          float d[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = v[ch][j] - mean;
          sq += ((d[0] * d[0] + d[1] * d[1]) + (d[2] * d[2] + d[3] * d[3])) + ((d[4] * d[4] + d[5] * d[5]) + (d[6] * d[6] + d[7] * d[7]));
        }
      }
      a = rsqrtf(group_sum<LPR>(sq) * invH + p.eps);
      nb = -mean * a;
    }
    if (!row_ok) continue;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c = (ch * LPR + sub) * 8;
      if (c >= H) continue;
      if (p.x0_f32) store8_f32(p.x0_f32 + t * H + c, v[ch]);
      if (p.x0) *reinterpret_cast<uint4*>(p.x0 + t * p.ldx + c) = pack8(v[ch]);
      if (has_ln) {
        float w[8], b[8], y[8];
        load8_f32(p.ln_w + c, w);
        load8_f32(p.ln_b + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = fmaf(fmaf(v[ch][j], a, nb), w[j], b[j]);
        *reinterpret_cast<uint4*>(p.q + t * p.ldx + c) = pack8(y);
      }
    }
    if (has_ln && p.stats && sub == 0) *reinterpret_cast<float2*>(p.stats + 2 * t) = make_float2(mean, a);
  }
}

// ---------------------------------------------------------------------------------------------
struct LnFwdParams {
  const bf16* x; int ldx;
  const float* w; const float* b; float eps;
  bf16* y_bf16; float* y_f32; int ldy;
  float* stats;
  int64_t T; int H;
  int64_t row_stride;   // process rows t*row_stride + row_offset (used to normalise only the last position)
  int64_t row_offset;
  const int* row_index; // optional: input row of output row t (packed layout: the row that holds each sequence's last position)
  const int* rows_dev;  // optional: device row count (packed layout)
};

// p.H is the number of VALID columns (LayerNorm width); rows are processed in 8-column chunks up to Hc = roundup(H, 8),
// columns in [H, Hc) are read as zero-padding and written as zeros.
template <int LPR, int CH, bool RAGGED>
__global__ void __launch_bounds__(256) ln_fwd_kernel(LnFwdParams p) {
  constexpr int RPW = 32 / LPR, UN = 2;
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  const int Hc = (p.H + 7) & ~7;
  constexpr bool ragged = RAGGED;
  pdl_prologue_done();
  if (p.rows_dev) p.T = min(p.T, (int64_t)__ldg(p.rows_dev));
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t0 = warp * RPW * UN; t0 < p.T; t0 += nwarps * RPW * UN) {
    uint4 raw[UN][CH];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t t = t0 + u * RPW + grp;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        const int c = (ch * LPR + sub) * 8;
        raw[u][ch] = make_uint4(0, 0, 0, 0);
        if (t < p.T && c < Hc)
          raw[u][ch] = __ldg(reinterpret_cast<const uint4*>(
              p.x + (p.row_index ? (int64_t)__ldg(p.row_index + t) : t * p.row_stride + p.row_offset) * p.ldx + c));
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int64_t t = t0 + u * RPW + grp;
      float v[CH][8];
      float sum = 0.f;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        unpack8(raw[u][ch], v[ch]);
        const int c = (ch * LPR + sub) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (ragged && c + j >= p.H) v[ch][j] = 0.f;
          sum += v[ch][j];
        }
      }
      const float mean = group_sum<LPR>(sum) / p.H;
      float sq = 0.f;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        const int c = (ch * LPR + sub) * 8;
        if (c < Hc) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = v[ch][j] - mean; sq += (!ragged || c + j < p.H) ? d * d : 0.f; }
        }
      }
      const float rstd = rsqrtf(group_sum<LPR>(sq) / p.H + p.eps);
      if (t >= p.T) continue;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        const int c = (ch * LPR + sub) * 8;
        if (c >= Hc) continue;
        float y[8], w[8], b[8];
        if (!ragged) { load8_f32(p.w + c, w); load8_f32(p.b + c, b); }
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { w[j] = c + j < p.H ? __ldg(p.w + c + j) : 0.f; b[j] = c + j < p.H ? __ldg(p.b + c + j) : 0.f; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (!ragged || c + j < p.H) ? (v[ch][j] - mean) * rstd * w[j] + b[j] : 0.f;
        if (p.y_bf16) *reinterpret_cast<uint4*>(p.y_bf16 + t * p.ldy + c) = pack8(y);
        if (p.y_f32) store8_f32(p.y_f32 + t * p.ldy + c, y);
      }
      if (p.stats && sub == 0) *reinterpret_cast<float2*>(p.stats + 2 * t) = make_float2(mean, rstd);
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct LnBwdParams {
  const bf16* dy_bf16; const float* dy_f32; int lddy;
  const bf16* x; int ldx;            // LN input
  const float* stats;                // (T, 2)
  const float* w;
  const bf16* add; int ldadd;        // optional: dx += add
  const int64_t* row_ids;            // optional: dx *= (row_ids[t] != 0)
  bf16* dx; int lddx;
  float* dw; float* db;              // (H) accumulated with red.add
  int64_t T; int H;
  const int* rows_dev;               // optional: device row count (packed layout)
};

// Lean register layout: the per-lane column accumulators (dw, db: 2 x CH x 8) are the only persistent state; the LN
// weight is re-read from shared memory, dy / x / add stay packed (bf16) until used, and g = dy * w and
// xh = (x - mean) * rstd are recomputed in the output pass instead of being kept.
template <int LPR, int CH, bool F32DY, bool RAGGED>
__global__ void __launch_bounds__(256, 2) ln_bwd_kernel(LnBwdParams p) {
  __shared__ float sdw[MAXW], sdb[MAXW];
  __shared__ __align__(16) float sw[MAXW];
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  const int Hc = (p.H + 7) & ~7;                         // p.H = valid columns; [H, Hc) is zero padding
  constexpr bool ragged = RAGGED;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  pdl_prologue_done();
  if (p.rows_dev) p.T = min(p.T, (int64_t)__ldg(p.rows_dev));
  for (int i = threadIdx.x; i < MAXW; i += blockDim.x) { sdw[i] = sdb[i] = 0.f; sw[i] = i < p.H ? __ldg(p.w + i) : 0.f; }
  __syncthreads();
  float adw[CH][8], adb[CH][8];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int j = 0; j < 8; ++j) adw[ch][j] = adb[ch][j] = 0.f;
  const float invH = 1.f / p.H;
  for (int64_t t0 = warp * RPW; t0 < p.T; t0 += nwarps * RPW) {
    const int64_t t = t0 + grp;
    const bool ok = t < p.T;
    uint4 xr[CH], ar[CH], dr[CH];
    float df[F32DY ? CH : 1][8];
    const float2 st = ok ? __ldg(reinterpret_cast<const float2*>(p.stats + 2 * t)) : make_float2(0.f, 0.f);
    const float msk = (ok && p.row_ids) ? ((__ldg(p.row_ids + t) != 0) ? 1.f : 0.f) : 1.f;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c = (ch * LPR + sub) * 8;
      xr[ch] = ar[ch] = dr[ch] = make_uint4(0, 0, 0, 0);
      if (F32DY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) df[ch][j] = 0.f;
      }
      if (ok && c < Hc) {
        if (F32DY) load8_f32(p.dy_f32 + t * p.lddy + c, df[ch]);
        else dr[ch] = __ldg(reinterpret_cast<const uint4*>(p.dy_bf16 + t * p.lddy + c));
        xr[ch] = __ldg(reinterpret_cast<const uint4*>(p.x + t * p.ldx + c));
        if (p.add) ar[ch] = __ldg(reinterpret_cast<const uint4*>(p.add + t * p.ldadd + c));
      }
    }
    const float mean = st.x, rstd = st.y;
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c = (ch * LPR + sub) * 8;
      const bool cok = c < Hc && ok;
      float xv[8], dv[8];
      unpack8(xr[ch], xv);
      if (F32DY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dv[j] = df[ch][j];
      } else {
        unpack8(dr[ch], dv);
      }
      const float4 w0 = *reinterpret_cast<const float4*>(sw + (cok ? c : 0)), w1 = *reinterpret_cast<const float4*>(sw + (cok ? c : 0) + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (cok && (!ragged || c + j < p.H)) ? (xv[j] - mean) * rstd : 0.f;
        if (ragged && c + j >= p.H) dv[j] = 0.f;
        const float g = dv[j] * wv[j];
        adw[ch][j] = fmaf(dv[j], xh, adw[ch][j]);
        adb[ch][j] += dv[j];
        sg += g;
        sgx = fmaf(g, xh, sgx);
      }
    }
    sg = group_sum<LPR>(sg) * invH;
    sgx = group_sum<LPR>(sgx) * invH;
    if (!ok) continue;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const int c = (ch * LPR + sub) * 8;
      if (c >= Hc) continue;
      float xv[8], dv[8], av[8], o[8];
      unpack8(xr[ch], xv);
      unpack8(ar[ch], av);
      if (F32DY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dv[j] = df[ch][j];
      } else {
        unpack8(dr[ch], dv);
      }
      const float4 w0 = *reinterpret_cast<const float4*>(sw + c), w1 = *reinterpret_cast<const float4*>(sw + c + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mean) * rstd;
        o[j] = (rstd * (dv[j] * wv[j] - sg - xh * sgx) + av[j]) * msk;
        if (ragged && c + j >= p.H) o[j] = 0.f;
      }
      *reinterpret_cast<uint4*>(p.dx + t * p.lddx + c) = pack8(o);
    }
  }
  // column partials: first across the row groups of the warp (lanes with equal `sub` hold the same columns), then one
  // shared-memory atomic per column per warp, then one red.add per column per block
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        adw[ch][j] += __shfl_xor_sync(0xffffffffu, adw[ch][j], o);
        adb[ch][j] += __shfl_xor_sync(0xffffffffu, adb[ch][j], o);
      }
    }
    const int c = (ch * LPR + sub) * 8;
    if (grp == 0 && c < Hc) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(&sdw[c + j], adw[ch][j]); atomicAdd(&sdb[c + j], adb[ch][j]); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.H; i += blockDim.x) { red_add_f32(p.dw + i, sdw[i]); red_add_f32(p.db + i, sdb[i]); }
}

// ---------------------------------------------------------------------------------------------
// Lean variants for widths that are multiples of 8 (every width the engine produces: ragged widths are zero-padded to 16).
// Same arithmetic as the generic kernels above; what changes is the instruction stream.  ncu on the generic forward kernel
// at C2 (204800 x 80, L2-resident): 760 warp instructions per 16 rows, half of the issue slots busy at 22 % occupancy --
// instruction-bound, not memory-bound.  Here only the LAST chunk of a lane can lie beyond H (the dispatch guarantees
// H > LPR*8*(CH-1)), so its loads are redirected to column 0 and zeroed with selects and only the stores are predicated: no
// divergent branches; the row index is clamped instead of tested; (x - mean) * rstd * w + b is two FMAs.
template <int LPR, int CH>
__global__ void __launch_bounds__(256, 4) ln_fwd_vec_kernel(LnFwdParams p) {
  constexpr int RPW = 32 / LPR, LAST = CH - 1;
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  pdl_prologue_done();
  const float invH = 1.f / (float)p.H;
  const bool last_ok = (LAST * LPR + sub) * 8 < p.H;
  int col[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) col[ch] = (ch * LPR + sub) * 8;
  if (!last_ok) col[LAST] = 0;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5) * RPW;
  if (p.rows_dev) p.T = min(p.T, (int64_t)__ldg(p.rows_dev));
  for (int64_t t = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + grp; t - grp < p.T; t += stride) {
    const bool ok = t < p.T;
    const int64_t tt = ok ? t : p.T - 1;
    const bf16* xrow = p.x + (p.row_index ? (int64_t)__ldg(p.row_index + tt) : tt * p.row_stride + p.row_offset) * p.ldx;
    uint4 raw[CH];
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) raw[ch] = __ldg(reinterpret_cast<const uint4*>(xrow + col[ch]));
    if (!last_ok) raw[LAST] = make_uint4(0, 0, 0, 0);
    float v[CH][8];
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      unpack8(raw[ch], v[ch]);
      sum += ((v[ch][0] + v[ch][1]) + (v[ch][2] + v[ch][3])) + ((v[ch][4] + v[ch][5]) + (v[ch][6] + v[ch][7]));
    }
    const float mean = group_sum<LPR>(sum) * invH;
    float sq = 0.f;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      float d[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = v[ch][j] - mean;
      const float s8 = ((d[0] * d[0] + d[1] * d[1]) + (d[2] * d[2] + d[3] * d[3])) + ((d[4] * d[4] + d[5] * d[5]) + (d[6] * d[6] + d[7] * d[7]));
      sq += (ch == LAST && !last_ok) ? 0.f : s8;
    }
    const float a = rsqrtf(group_sum<LPR>(sq) * invH + p.eps), nb = -mean * a;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      float y[8], w[8], b[8];
      load8_f32(p.w + col[ch], w);
      load8_f32(p.b + col[ch], b);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaf(fmaf(v[ch][j], a, nb), w[j], b[j]);
      const bool st = ok && (ch != LAST || last_ok);
      if (p.y_bf16 && st) *reinterpret_cast<uint4*>(p.y_bf16 + t * p.ldy + col[ch]) = pack8(y);
      if (p.y_f32 && st) store8_f32(p.y_f32 + t * p.ldy + col[ch], y);
    }
    if (p.stats && sub == 0 && ok) *reinterpret_cast<float2*>(p.stats + 2 * t) = make_float2(mean, a);
  }
}

// Backward, same treatment.  Zeroing dy of an out-of-range chunk makes every one of its contributions (dw, db, the two row
// sums) vanish, so no further masking is needed; the row mask is applied to the packed output.
template <int LPR, int CH, bool F32DY, bool HAS_ADD, int TPB, int BPS>
__global__ void __launch_bounds__(TPB, BPS) ln_bwd_vec_kernel(LnBwdParams p) {
  __shared__ float sdw[MAXW], sdb[MAXW];
  __shared__ __align__(16) float sw[MAXW];
  constexpr int RPW = 32 / LPR, LAST = CH - 1;
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  pdl_prologue_done();
  if (p.rows_dev) p.T = min(p.T, (int64_t)__ldg(p.rows_dev));
  for (int i = threadIdx.x; i < MAXW; i += blockDim.x) { sdw[i] = sdb[i] = 0.f; sw[i] = i < p.H ? __ldg(p.w + i) : 0.f; }
  __syncthreads();
  float adw[CH][8], adb[CH][8];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int j = 0; j < 8; ++j) adw[ch][j] = adb[ch][j] = 0.f;
  const float invH = 1.f / (float)p.H;
  const bool last_ok = (LAST * LPR + sub) * 8 < p.H;
  int col[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) col[ch] = (ch * LPR + sub) * 8;
  if (!last_ok) col[LAST] = 0;
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5) * RPW;
  for (int64_t t = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + grp; t - grp < p.T; t += stride) {
    const bool ok = t < p.T;
    const int64_t tt = ok ? t : p.T - 1;
    uint4 xr[CH], ar[HAS_ADD ? CH : 1], dr[F32DY ? 1 : CH];
    float df[F32DY ? CH : 1][8];
    const float2 st = __ldg(reinterpret_cast<const float2*>(p.stats + 2 * tt));
    const bool keep = p.row_ids ? (__ldg(p.row_ids + tt) != 0) : true;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const bool cok = ok && (ch != LAST || last_ok);
      if (F32DY) {
        load8_f32(p.dy_f32 + tt * p.lddy + col[ch], df[ch]);
        if (!cok) {
#pragma unroll
          for (int j = 0; j < 8; ++j) df[ch][j] = 0.f;
        }
      } else {
        dr[ch] = __ldg(reinterpret_cast<const uint4*>(p.dy_bf16 + tt * p.lddy + col[ch]));
        if (!cok) dr[ch] = make_uint4(0, 0, 0, 0);
      }
      xr[ch] = __ldg(reinterpret_cast<const uint4*>(p.x + tt * p.ldx + col[ch]));
      if (HAS_ADD) ar[ch] = __ldg(reinterpret_cast<const uint4*>(p.add + tt * p.ldadd + col[ch]));
    }
    const float rstd = st.y, nb = -st.x * st.y;
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      float xv[8], dv[8];
      unpack8(xr[ch], xv);
      if (F32DY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dv[j] = df[ch][j];
      } else {
        unpack8(dr[ch], dv);
      }
      const float4 w0 = *reinterpret_cast<const float4*>(sw + col[ch]), w1 = *reinterpret_cast<const float4*>(sw + col[ch] + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(xv[j], rstd, nb);
        const float g = dv[j] * wv[j];
        adw[ch][j] = fmaf(dv[j], xh, adw[ch][j]);
        adb[ch][j] += dv[j];
        sg += g;
        sgx = fmaf(g, xh, sgx);
      }
    }
    sg = group_sum<LPR>(sg) * invH;
    sgx = group_sum<LPR>(sgx) * invH;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      float xv[8], dv[8], av[8], o[8];
      unpack8(xr[ch], xv);
      if (HAS_ADD) unpack8(ar[ch], av);
      if (F32DY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dv[j] = df[ch][j];
      } else {
        unpack8(dr[ch], dv);
      }
      const float4 w0 = *reinterpret_cast<const float4*>(sw + col[ch]), w1 = *reinterpret_cast<const float4*>(sw + col[ch] + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(xv[j], rstd, nb);
        const float u = fmaf(-sgx, xh, fmaf(dv[j], wv[j], -sg));
        o[j] = HAS_ADD ? fmaf(rstd, u, av[j]) : rstd * u;
      }
      uint4 pk = pack8(o);
      if (!keep) pk = make_uint4(0, 0, 0, 0);
      if (ok && (ch != LAST || last_ok)) *reinterpret_cast<uint4*>(p.dx + t * p.lddx + col[ch]) = pk;
    }
  }
  // column partials: first across the row groups of the warp (lanes with equal `sub` hold the same columns), then one
  // shared-memory atomic per column per warp, then one red.add per column per block
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        adw[ch][j] += __shfl_xor_sync(0xffffffffu, adw[ch][j], o);
        adb[ch][j] += __shfl_xor_sync(0xffffffffu, adb[ch][j], o);
      }
    }
    if (grp == 0 && (ch != LAST || last_ok)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(&sdw[col[ch] + j], adw[ch][j]); atomicAdd(&sdb[col[ch] + j], adb[ch][j]); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.H; i += blockDim.x) { red_add_f32(p.dw + i, sdw[i]); red_add_f32(p.db + i, sdb[i]); }
}

// dispatch on the row width: lanes per row x chunks per lane (8 features each)
// backward keeps more per-lane state (column accumulators): one chunk per lane up to 256 columns
#define SRFRD_ROW_DISPATCH_BWD(H, CALL)                  \
  do {                                                   \
    if ((H) <= 64) { CALL(8, 1); }                       \
    else if ((H) <= 128) { CALL(16, 1); }                \
    else if ((H) <= 256) { CALL(32, 1); }                \
    else { CALL(32, 2); }                                \
  } while (0)
// (the capacity LPR * CH * 8 closest above H wins: H = 80 runs as 4 lanes x 3 chunks = 96 columns, 83 % of the
//  lanes busy and two shuffle steps per reduction, instead of 16 x 1 = 128 columns at 62 %)
#define SRFRD_ROW_DISPATCH(H, CALL)                      \
  do {                                                   \
    if ((H) <= 32) { CALL(4, 1); }                       \
    else if ((H) <= 64) { CALL(8, 1); }                  \
    else if ((H) <= 96) { CALL(4, 3); }                  \
    else if ((H) <= 128) { CALL(16, 1); }                \
    else if ((H) <= 192) { CALL(8, 3); }                 \
    else if ((H) <= 256) { CALL(32, 1); }                \
    else if ((H) <= 384) { CALL(16, 3); }                \
    else { CALL(32, 2); }                                \
  } while (0)

// ---------------------------------------------------------------------------------------------
// out[n] += sum_m X[m, n]  (bias gradients; positional-table gradient via the (B, L*H) view)
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* X, int64_t M, int N, int64_t ld, float* out,
                                                     int rows_per_block) {
  // blockDim = (TX, TY): thread x handles column pairs, y strides rows
  extern __shared__ float sm[];
  pdl_prologue_done();
  const int TX = blockDim.x, TY = blockDim.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  for (int c0 = blockIdx.y * TX * 2; c0 < N; c0 += gridDim.y * TX * 2) {
    const int c = c0 + threadIdx.x * 2;
    float a0 = 0.f, a1 = 0.f;
    if (c < N) {
      for (int64_t r = r0 + threadIdx.y; r < r1; r += TY) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(X + r * ld + c);
        a0 += __low2float(v);
        a1 += __high2float(v);
      }
    }
    sm[(threadIdx.y * TX + threadIdx.x) * 2] = a0;
    sm[(threadIdx.y * TX + threadIdx.x) * 2 + 1] = a1;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
      for (int y = 1; y < TY; ++y) { a0 += sm[(y * TX + threadIdx.x) * 2]; a1 += sm[(y * TX + threadIdx.x) * 2 + 1]; }
      red_add_f32(out + c, a0);
      red_add_f32(out + c + 1, a1);
    }
    __syncthreads();
  }
}

// dst[r, c] = bf16(src[r, c]), dstT[c, r] = bf16(src[r, c]) for a table of small weight matrices
__global__ void cast_weights_kernel(const srfrd_cast_desc_t* descs, int n) {
  pdl_prologue_done();
  const srfrd_cast_desc_t d = descs[blockIdx.x];
  const int total = d.rows * d.cols;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const int r = i / d.cols, c = i % d.cols;
    if (d.dst_is_f32) {
      reinterpret_cast<float*>(d.dst)[(size_t)r * d.dst_ld + c] = d.src[(size_t)r * d.src_ld + c];
      continue;
    }
    const bf16 v = f2bf(d.src[(size_t)r * d.src_ld + c]);
    if (d.dst) reinterpret_cast<bf16*>(d.dst)[(size_t)r * d.dst_ld + c] = v;
    if (d.dst_t) reinterpret_cast<bf16*>(d.dst_t)[(size_t)c * d.dst_t_ld + r] = v;
  }
}

// SRFU_B/F/R.get_Labels (SRFR_model.py:546-570): one warp per sequence
__global__ void srfu_labels_kernel(const int64_t* fake_ids, int64_t B, int L, int kind, int64_t* labels) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int nf = 0, nr = 0;
  for (int l = lane; l < L; l += 32) {
    const int64_t f = fake_ids[b * L + l];
    nf += f == 1;
    nr += f == 2;
  }
  for (int o = 16; o > 0; o >>= 1) { nf += __shfl_xor_sync(0xffffffffu, nf, o); nr += __shfl_xor_sync(0xffffffffu, nr, o); }
  if (lane == 0) {
    int64_t lab;
    if (kind == 0) lab = nf < nr ? 1 : 2;                      // round(sign(nf-nr)*0.5+1.5), tie -> 2
    else if (kind == 3) lab = nf > nr ? 2 : 1;                 // SRFRN.predict: (sign(..)*0.5+1.5).int(), tie -> 1 (:244)
    else if (kind == 1) lab = nf;
    else {
      const float r = __fmul_rn(__fdiv_rn((float)nf, (float)(nf + nr)), 10.f);
      lab = (nf + nr) == 0 ? 0 : (int64_t)floorf(r);           // reference: 0/0 -> NaN -> undefined id
    }
    labels[b] = lab;
  }
}

__global__ void f32_to_bf16_rows_kernel(const float* src, int64_t src_ld, const int64_t* row_index, bf16* hi, bf16* lo,
                                        int64_t rows, int cols, int dst_ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int64_t r = i / cols; const int c = (int)(i % cols);
  const int64_t sr = row_index ? row_index[r] : r;
  const float v = src[sr * src_ld + c];
  const bf16 h = f2bf(v);
  hi[r * dst_ld + c] = h;
  if (lo) lo[r * dst_ld + c] = f2bf(v - bf2f(h));
}

__global__ void dropout_apply_kernel(const bf16* x, int ldx, bf16* out, int ldo, int64_t M, int N, uint64_t seed,
                                     uint32_t thresh, uint32_t stream_id, float scale, const float* step,
                                     const int* rows_dev) {
  pdl_prologue_done();
  if (rows_dev) M = min(M, (int64_t)__ldg(rows_dev));
  seed = mix_seed(seed, step);
  const int64_t n2 = M * (N / 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (N / 2); const int c = (int)(i % (N / 2)) * 2;
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * ldx + c));
    const float a = dropout_keep(seed, stream_id, (uint64_t)r * N + c, thresh) ? v.x * scale : 0.f;
    const float b = dropout_keep(seed, stream_id, (uint64_t)r * N + c + 1, thresh) ? v.y * scale : 0.f;
    *reinterpret_cast<uint32_t*>(out + r * ldo + c) = pack_bf16x2(a, b);
  }
}

static int grid_for_rows(int64_t T, int rows_per_block, int max_blocks_per_sm) {
  int64_t blocks = (T + rows_per_block - 1) / rows_per_block;
  const int64_t cap = (int64_t)num_sms() * max_blocks_per_sm;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace srfrd

using namespace srfrd;

static int embed_ln_fwd_impl(const float* item_table, int64_t n_rows, int D, const float* pos_table,
                             const float* aux_table, int64_t n_aux, int F, int mode, const int64_t* seq,
                             const int64_t* aux_ids, int64_t B, int L, float item_scale, const float* ln_w,
                             const float* ln_b, float eps, void* x0_bf16, float* x0_f32, void* q_bf16,
                             float* stats, int ldx, float drop_p, uint64_t drop_seed, uint32_t drop_stream,
                             const float* drop_step, const int* row_tok, const int* rows_dev, int64_t cap_rows,
                             void* stream) {
  SRFRD_REQUIRE(item_table && pos_table && seq, "embed_ln_fwd: null input");
  SRFRD_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "embed_ln_fwd: dropout p must be in [0, 1)");
  SRFRD_REQUIRE(mode >= 0 && mode <= 2, "embed_ln_fwd: mode must be 0 (none), 1 (concat fake) or 2 (add user label)");
  SRFRD_REQUIRE(mode == 0 || aux_table, "embed_ln_fwd: aux_table required for mode %d", mode);
  SRFRD_REQUIRE(mode != 2 || aux_ids, "embed_ln_fwd: per-sequence labels required for mode 2");
  const int H = D + (mode == 1 ? F : 0);
  SRFRD_REQUIRE(H <= MAXW, "embed_ln_fwd: width %d > %d unsupported", H, MAXW);
  const bool vec = D % 8 == 0 && H % 8 == 0;
  SRFRD_REQUIRE(!ln_w || (ln_b && q_bf16), "embed_ln_fwd: LN needs weight, bias and an output");
  SRFRD_REQUIRE((!x0_bf16 && !q_bf16) || (ldx >= ((H + 7) & ~7) && ldx % 8 == 0), "embed_ln_fwd: bad ldx %d", ldx);
  (void)n_rows; (void)n_aux;
  if (B * L == 0) return 0;
  SRFRD_REQUIRE(!row_tok || (vec && rows_dev && cap_rows > 0 && !x0_f32),
                "embed_ln_fwd_packed: needs widths that are multiples of 8, the device row count and a capacity");
  EmbedParams p;
  p.item_table = item_table; p.pos_table = pos_table; p.aux_table = aux_table; p.seq = seq; p.aux_ids = aux_ids;
  p.row_tok = row_tok; p.rows_dev = rows_dev;
  p.T = row_tok ? cap_rows : B * L; p.L = L; p.D = D; p.F = F; p.mode = mode; p.item_scale = item_scale;
  p.ln_w = ln_w; p.ln_b = ln_b; p.eps = eps; p.x0 = (bf16*)x0_bf16; p.x0_f32 = x0_f32; p.q = (bf16*)q_bf16;
  p.stats = stats; p.ldx = ldx;
  p.drop_seed = drop_seed; p.drop_stream = drop_stream; p.drop_step = drop_step;
  p.drop_thresh = drop_p > 0.f ? (uint32_t)((double)drop_p * 4294967296.0) : 0;
  p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#define CALL(LPR, CH)                                                                                       \
  do {                                                                                                      \
    const int grid = grid_for_rows(p.T, 8 * (32 / LPR), 8);                                                 \
    if (vec) {                                                                                              \
      auto* kern = !p.drop_thresh ? (p.mode == 0 ? embed_ln_vec_kernel<LPR, CH, 0, false>                   \
                                    : p.mode == 1 ? embed_ln_vec_kernel<LPR, CH, 1, false>                  \
                                                  : embed_ln_vec_kernel<LPR, CH, 2, false>)                 \
                                  : (p.mode == 0 ? embed_ln_vec_kernel<LPR, CH, 0, true>                    \
                                    : p.mode == 1 ? embed_ln_vec_kernel<LPR, CH, 1, true>                   \
                                                  : embed_ln_vec_kernel<LPR, CH, 2, true>);                 \
      SRFRD_CUDA(launch_pdl(kern, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p));                      \
    }                                                                                                       \
    else SRFRD_CUDA(launch_pdl(embed_ln_kernel<LPR, CH, false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p));    \
  } while (0)
  SRFRD_ROW_DISPATCH(H, CALL);
#undef CALL
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_embed_ln_fwd(const float* item_table, int64_t n_rows, int D, const float* pos_table,
                                  const float* aux_table, int64_t n_aux, int F, int mode, const int64_t* seq,
                                  const int64_t* aux_ids, int64_t B, int L, float item_scale, const float* ln_w,
                                  const float* ln_b, float eps, void* x0_bf16, float* x0_f32, void* q_bf16,
                                  float* stats, int ldx, float drop_p, uint64_t drop_seed, uint32_t drop_stream,
                                  const float* drop_step, void* stream) {
  return embed_ln_fwd_impl(item_table, n_rows, D, pos_table, aux_table, n_aux, F, mode, seq, aux_ids, B, L, item_scale, ln_w,
                           ln_b, eps, x0_bf16, x0_f32, q_bf16, stats, ldx, drop_p, drop_seed, drop_stream, drop_step, nullptr,
                           nullptr, 0, stream);
}

extern "C" int srfrd_embed_ln_fwd_packed(const float* item_table, int64_t n_rows, int D, const float* pos_table,
                                         const float* aux_table, int64_t n_aux, int F, int mode, const int64_t* seq,
                                         const int64_t* aux_ids, int64_t B, int L, float item_scale, const float* ln_w,
                                         const float* ln_b, float eps, void* x0_bf16, void* q_bf16, float* stats, int ldx,
                                         float drop_p, uint64_t drop_seed, uint32_t drop_stream, const float* drop_step,
                                         const int* row_tok, const int* rows_dev, int64_t cap_rows, void* stream) {
  SRFRD_REQUIRE(row_tok && rows_dev, "embed_ln_fwd_packed: null row map");
  return embed_ln_fwd_impl(item_table, n_rows, D, pos_table, aux_table, n_aux, F, mode, seq, aux_ids, B, L, item_scale, ln_w,
                           ln_b, eps, x0_bf16, nullptr, q_bf16, stats, ldx, drop_p, drop_seed, drop_stream, drop_step, row_tok,
                           rows_dev, cap_rows, stream);
}

static int layernorm_fwd_impl(const void* x, int ldx, const float* w, const float* b, float eps, void* y_bf16,
                              float* y_f32, int ldy, float* stats, int64_t T, int H, int64_t row_stride,
                              int64_t row_offset, const int* row_index, const int* rows_dev, void* stream) {
  SRFRD_REQUIRE(x && w && b && (y_bf16 || y_f32), "layernorm_fwd: null pointer");
  SRFRD_REQUIRE(H <= MAXW && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= ((H + 7) & ~7) && ldy >= ((H + 7) & ~7),
                "layernorm_fwd: width %d / ld (%d, %d) unsupported (rows are padded to a multiple of 8 columns)", H, ldx, ldy);
  if (T == 0) return 0;
  LnFwdParams p;
  p.x = (const bf16*)x; p.ldx = ldx; p.w = w; p.b = b; p.eps = eps; p.y_bf16 = (bf16*)y_bf16; p.y_f32 = y_f32;
  p.ldy = ldy; p.stats = stats; p.T = T; p.H = H; p.row_stride = row_stride; p.row_offset = row_offset;
  p.row_index = row_index; p.rows_dev = rows_dev;
#define CALL(LPR, CH)                                                                                                   \
  do {                                                                                                                  \
    const dim3 grid(grid_for_rows(T, 16 * (32 / LPR), 8));                                                              \
    if (H % 8) SRFRD_CUDA(launch_pdl(ln_fwd_kernel<LPR, CH, true>, grid, dim3(256), 0, (cudaStream_t)stream, p));       \
    else SRFRD_CUDA(launch_pdl(ln_fwd_vec_kernel<LPR, CH>, dim3(grid_for_rows(T, 8 * (32 / LPR), 4)), dim3(256), 0,     \
                               (cudaStream_t)stream, p));                                                               \
  } while (0)
  SRFRD_ROW_DISPATCH(H, CALL);
#undef CALL
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_layernorm_fwd(const void* x, int ldx, const float* w, const float* b, float eps, void* y_bf16,
                                   float* y_f32, int ldy, float* stats, int64_t T, int H, int64_t row_stride,
                                   int64_t row_offset, void* stream) {
  return layernorm_fwd_impl(x, ldx, w, b, eps, y_bf16, y_f32, ldy, stats, T, H, row_stride, row_offset, nullptr, row_limit(),
                            stream);
}

extern "C" int srfrd_layernorm_fwd_rows(const void* x, int ldx, const float* w, const float* b, float eps, float* y_f32,
                                        int ldy, const int* row_index, int64_t n, int H, void* stream) {
  SRFRD_REQUIRE(row_index, "layernorm_fwd_rows: null row index");
  return layernorm_fwd_impl(x, ldx, w, b, eps, nullptr, y_f32, ldy, nullptr, n, H, 1, 0, row_index, nullptr, stream);
}

extern "C" int srfrd_layernorm_bwd(const void* dy_bf16, const float* dy_f32, int lddy, const void* x, int ldx,
                                   const float* stats, const float* w, const void* add, int ldadd,
                                   const int64_t* row_ids, void* dx, int lddx, float* dw, float* db, int64_t T, int H,
                                   void* stream) {
  SRFRD_REQUIRE((dy_bf16 || dy_f32) && x && stats && w && dx && dw && db, "layernorm_bwd: null pointer");
  const int Hc8 = (H + 7) & ~7;
  SRFRD_REQUIRE(H <= MAXW - 8 && ldx % 8 == 0 && lddx % 8 == 0 && lddy % 8 == 0 && (!add || ldadd % 8 == 0) &&
                    ldx >= Hc8 && lddx >= Hc8 && lddy >= Hc8 && (!add || ldadd >= Hc8),
                "layernorm_bwd: width %d / ld unsupported (rows are padded to a multiple of 8 columns)", H);
  if (T == 0) return 0;
  LnBwdParams p;
  p.dy_bf16 = (const bf16*)dy_bf16; p.dy_f32 = dy_f32; p.lddy = lddy; p.x = (const bf16*)x; p.ldx = ldx;
  p.stats = stats; p.w = w; p.add = (const bf16*)add; p.ldadd = ldadd; p.row_ids = row_ids; p.dx = (bf16*)dx;
  p.lddx = lddx; p.dw = dw; p.db = db; p.T = T; p.H = H; p.rows_dev = row_limit();
#define CALL(LPR, CH)                                                                                             \
  do {                                                                                                            \
    const int grid = grid_for_rows(T, 8 * (32 / LPR) * 8, 2);                                                     \
    if (dy_f32 && (H % 8)) SRFRD_CUDA(launch_pdl(ln_bwd_kernel<LPR, CH, true, true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p));        \
    else if (H % 8) SRFRD_CUDA(launch_pdl(ln_bwd_kernel<LPR, CH, false, true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p));        \
    else {                                                                                                        \
      /* 128 threads x 3 blocks per SM: 168 registers per thread, no spills (256 x 2 caps at 128 and spills) */  \
      const int g3 = grid_for_rows(T, 4 * (32 / LPR) * 8, 3);                                                     \
      if (dy_f32 && add) SRFRD_CUDA(launch_pdl(ln_bwd_vec_kernel<LPR, CH, true, true, 128, 3>, dim3(g3), dim3(128), 0, (cudaStream_t)stream, p));   \
      else if (dy_f32) SRFRD_CUDA(launch_pdl(ln_bwd_vec_kernel<LPR, CH, true, false, 128, 3>, dim3(g3), dim3(128), 0, (cudaStream_t)stream, p));  \
      else if (add) SRFRD_CUDA(launch_pdl(ln_bwd_vec_kernel<LPR, CH, false, true, 128, 3>, dim3(g3), dim3(128), 0, (cudaStream_t)stream, p));     \
      else SRFRD_CUDA(launch_pdl(ln_bwd_vec_kernel<LPR, CH, false, false, 128, 3>, dim3(g3), dim3(128), 0, (cudaStream_t)stream, p));             \
    }                                                                                                             \
  } while (0)
  SRFRD_ROW_DISPATCH(H, CALL);
#undef CALL
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_colsum(const void* X, int64_t M, int N, int64_t ld, float* out, void* stream) {
  SRFRD_REQUIRE(X && out, "colsum: null pointer");
  SRFRD_REQUIRE(N % 2 == 0 && ld % 2 == 0, "colsum: N and ld must be even");
  if (M == 0 || N == 0) return 0;
  int tx = 32;
  while (tx * 2 < N && tx < 128) tx *= 2;
  const int ty = 256 / tx;
  const int gy = (N + tx * 2 - 1) / (tx * 2);
  int64_t gx = (int64_t)num_sms() * 8 / gy;
  if (gx < 1) gx = 1;
  int64_t rows_per_block = (M + gx - 1) / gx;
  if (rows_per_block < ty * 4) rows_per_block = ty * 4;
  gx = (M + rows_per_block - 1) / rows_per_block;
  dim3 grid((unsigned)gx, (unsigned)gy), block(tx, ty);
  SRFRD_CUDA(launch_pdl(colsum_kernel, grid, block, 256 * 2 * sizeof(float), (cudaStream_t)stream, (const bf16*)X, M, N, ld, out,
                        (int)rows_per_block));
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_cast_weights(const srfrd_cast_desc_t* descs_dev, int n, void* stream) {
  SRFRD_REQUIRE(descs_dev || n == 0, "cast_weights: null table");
  if (n == 0) return 0;
  SRFRD_CUDA(launch_pdl(cast_weights_kernel, dim3(n, 8), dim3(256), 0, (cudaStream_t)stream, descs_dev, n));
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_srfu_labels(const int64_t* fake_ids, int64_t B, int L, int kind, int64_t* labels, void* stream) {
  SRFRD_REQUIRE(fake_ids && labels, "srfu_labels: null pointer");
  SRFRD_REQUIRE(kind >= 0 && kind <= 3, "srfu_labels: kind must be 0 (B), 1 (F), 2 (R) or 3 (SRFRN.predict)");
  if (B == 0) return 0;
  srfu_labels_kernel<<<(unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream>>>(fake_ids, B, L, kind, labels);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_f32_to_bf16_split(const float* src, int64_t src_ld, const int64_t* row_index, void* hi, void* lo,
                                       int64_t rows, int cols, int dst_ld, void* stream) {
  SRFRD_REQUIRE(src && hi, "f32_to_bf16_split: null pointer");
  if (rows * cols == 0) return 0;
  const int64_t n = rows * cols;
  f32_to_bf16_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      src, src_ld, row_index, (bf16*)hi, (bf16*)lo, rows, cols, dst_ld);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_dropout_apply(const void* x, int ldx, void* out, int ldo, int64_t M, int N, float drop_p,
                                   uint64_t seed, uint32_t stream_id, const float* drop_step, void* stream) {
  SRFRD_REQUIRE(x && out, "dropout_apply: null pointer");
  SRFRD_REQUIRE(N % 2 == 0 && ldx % 2 == 0 && ldo % 2 == 0, "dropout_apply: widths must be even");
  SRFRD_REQUIRE(drop_p > 0.f && drop_p < 1.f, "dropout_apply: p must be in (0, 1)");
  if (M == 0) return 0;
  int64_t blocks = (M * (N / 2) + 255) / 256;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  SRFRD_CUDA(launch_pdl(dropout_apply_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, ldx,
                        (bf16*)out, ldo, M, N, seed, (uint32_t)((double)drop_p * 4294967296.0), stream_id,
                        1.f / (1.f - drop_p), drop_step, row_limit()));
  SRFRD_LAUNCH_CHECK();
  return 0;
}

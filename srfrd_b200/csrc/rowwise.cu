// HBM-bound row-wise kernels: K1 (gather + positional add + concat/add + pad mask + LayerNorm),
// LayerNorm forward/backward, column sums (bias gradients), fp32 -> bf16 weight shadows.
// One warp per token; lanes stride the feature dimension so every global access is coalesced.
#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

static constexpr int MAXC = 16;  // features per lane -> widths up to 512

struct EmbedParams {
  const float* item_table;   // (n_rows, D)
  const float* pos_table;    // (L, D)
  const float* aux_table;    // mode 1: fake_embed (3, F); mode 2: user_label_embed (labels, D)
  const int64_t* seq;        // (B, L)
  const int64_t* aux_ids;    // mode 1: (B, L) fake ids or null (-> all 0); mode 2: (B,) labels
  int64_t T;
  int L, D, F, mode;
  int64_t n_rows, n_aux;
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

__global__ void __launch_bounds__(256) embed_ln_kernel(EmbedParams p) {
  if (p.drop_thresh) p.drop_seed = mix_seed(p.drop_seed, p.drop_step);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int H = p.D + (p.mode == 1 ? p.F : 0);
  for (int64_t t = warp0; t < p.T; t += nwarps) {
    const int l = (int)(t % p.L);
    const int64_t id = __ldg(p.seq + t);
    const bool valid = id != 0;
    int64_t aid = 0;
    if (p.mode == 1) aid = p.aux_ids ? __ldg(p.aux_ids + t) : 0;
    if (p.mode == 2) aid = __ldg(p.aux_ids + t / p.L);
    float v[MAXC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      float x = 0.f;
      if (c < H && valid) {
        if (c < p.D) {
          // __fmul_rn/__fadd_rn: no FMA contraction, so the pre-LN tensor is bit-identical to
          // torch's  E[id] (* sqrt(d)) + P[l] (+ Ul[label])   (SRFR_model.py:22-25, :622-624, :419-422)
          x = __ldg(p.item_table + id * p.D + c);
          if (p.item_scale != 1.f) x = __fmul_rn(x, p.item_scale);
          x = __fadd_rn(x, __ldg(p.pos_table + (int64_t)l * p.D + c));
          if (p.mode == 2) x = __fadd_rn(x, __ldg(p.aux_table + aid * p.D + c));
        } else {
          x = __ldg(p.aux_table + aid * p.F + (c - p.D));
        }
        // SASRec.emb_dropout (SRFR_model.py:625): after the positional add, before the pad mask
        if (p.drop_thresh)
          x = dropout_keep(p.drop_seed, p.drop_stream, (uint64_t)t * H + c, p.drop_thresh) ? x * p.drop_scale : 0.f;
      }
      v[i] = x;
      sum += x;
    }
    if (p.x0_f32) {
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < H) p.x0_f32[t * H + c] = v[i];
      }
    }
    if (p.x0) {
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < H) p.x0[t * p.ldx + c] = f2bf(v[i]);
      }
    }
    if (p.ln_w) {
      const float mean = warp_sum(sum) / H;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < H) { const float d = v[i] - mean; sq += d * d; }
      }
      const float rstd = rsqrtf(warp_sum(sq) / H + p.eps);
#pragma unroll
      for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < H) p.q[t * p.ldx + c] = f2bf((v[i] - mean) * rstd * __ldg(p.ln_w + c) + __ldg(p.ln_b + c));
      }
      if (p.stats && lane == 0) { p.stats[2 * t] = mean; p.stats[2 * t + 1] = rstd; }
    }
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
};

__global__ void __launch_bounds__(256) ln_fwd_kernel(LnFwdParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = warp0; t < p.T; t += nwarps) {
    const int64_t r = t * p.row_stride + p.row_offset;
    float v[MAXC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < p.H ? bf2f(p.x[r * p.ldx + c]) : 0.f;
      sum += v[i];
    }
    const float mean = warp_sum(sum) / p.H;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < p.H) { const float d = v[i] - mean; sq += d * d; }
    }
    const float rstd = rsqrtf(warp_sum(sq) / p.H + p.eps);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < p.H) {
        const float y = (v[i] - mean) * rstd * __ldg(p.w + c) + __ldg(p.b + c);
        if (p.y_bf16) p.y_bf16[t * p.ldy + c] = f2bf(y);
        if (p.y_f32) p.y_f32[t * p.ldy + c] = y;
      }
    }
    if (p.stats && lane == 0) { p.stats[2 * t] = mean; p.stats[2 * t + 1] = rstd; }
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
};

__global__ void __launch_bounds__(256) ln_bwd_kernel(LnBwdParams p) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * nw + warp;
  const int64_t nwarps = (int64_t)gridDim.x * nw;
  float adw[MAXC], adb[MAXC];
#pragma unroll
  for (int i = 0; i < MAXC; ++i) adw[i] = adb[i] = 0.f;
  for (int64_t t = warp0; t < p.T; t += nwarps) {
    const float mean = __ldg(p.stats + 2 * t), rstd = __ldg(p.stats + 2 * t + 1);
    float g[MAXC], xh[MAXC];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      g[i] = xh[i] = 0.f;
      if (c < p.H) {
        const float dy = p.dy_bf16 ? bf2f(p.dy_bf16[t * p.lddy + c]) : p.dy_f32[t * p.lddy + c];
        xh[i] = (bf2f(p.x[t * p.ldx + c]) - mean) * rstd;
        g[i] = dy * __ldg(p.w + c);
        adw[i] += dy * xh[i];
        adb[i] += dy;
        sg += g[i];
        sgx += g[i] * xh[i];
      }
    }
    sg = warp_sum(sg) / p.H;
    sgx = warp_sum(sgx) / p.H;
    float m = 1.f;
    if (p.row_ids) m = (__ldg(p.row_ids + t) != 0) ? 1.f : 0.f;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < p.H) {
        float dx = rstd * (g[i] - sg - xh[i] * sgx);
        if (p.add) dx += bf2f(p.add[t * p.ldadd + c]);
        p.dx[t * p.lddx + c] = f2bf(dx * m);
      }
    }
  }
  // block reduction of the per-lane column partials, then one red.add per column per block
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    if (32 * i >= p.H) break;
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
      red[warp][lane] = pass ? adb[i] : adw[i];
      __syncthreads();
      if (warp == 0) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += red[w][lane];
        const int c = lane + 32 * i;
        if (c < p.H) red_add_f32((pass ? p.db : p.dw) + c, s);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// out[n] += sum_m X[m, n]  (bias gradients; positional-table gradient via the (B, L*H) view)
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* X, int64_t M, int N, int64_t ld, float* out,
                                                     int rows_per_block) {
  // blockDim = (TX, TY): thread x handles column pairs, y strides rows
  extern __shared__ float sm[];
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
  const srfrd_cast_desc_t d = descs[blockIdx.x];
  const int total = d.rows * d.cols;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    const int r = i / d.cols, c = i % d.cols;
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
                                     uint32_t thresh, uint32_t stream_id, float scale, const float* step) {
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

static int grid_for_warps(int64_t T, int warps_per_block) {
  int64_t blocks = (T + warps_per_block - 1) / warps_per_block;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace srfrd

using namespace srfrd;

extern "C" int srfrd_embed_ln_fwd(const float* item_table, int64_t n_rows, int D, const float* pos_table,
                                  const float* aux_table, int64_t n_aux, int F, int mode, const int64_t* seq,
                                  const int64_t* aux_ids, int64_t B, int L, float item_scale, const float* ln_w,
                                  const float* ln_b, float eps, void* x0_bf16, float* x0_f32, void* q_bf16,
                                  float* stats, int ldx, float drop_p, uint64_t drop_seed, uint32_t drop_stream,
                                  const float* drop_step, void* stream) {
  SRFRD_REQUIRE(item_table && pos_table && seq, "embed_ln_fwd: null input");
  SRFRD_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "embed_ln_fwd: dropout p must be in [0, 1)");
  SRFRD_REQUIRE(mode >= 0 && mode <= 2, "embed_ln_fwd: mode must be 0 (none), 1 (concat fake) or 2 (add user label)");
  SRFRD_REQUIRE(mode == 0 || aux_table, "embed_ln_fwd: aux_table required for mode %d", mode);
  SRFRD_REQUIRE(mode != 2 || aux_ids, "embed_ln_fwd: per-sequence labels required for mode 2");
  const int H = D + (mode == 1 ? F : 0);
  SRFRD_REQUIRE(H <= 32 * MAXC, "embed_ln_fwd: width %d > %d unsupported", H, 32 * MAXC);
  SRFRD_REQUIRE(!ln_w || (ln_b && q_bf16), "embed_ln_fwd: LN needs weight, bias and an output");
  SRFRD_REQUIRE((!x0_bf16 && !q_bf16) || ldx >= H, "embed_ln_fwd: ldx < H");
  if (B * L == 0) return 0;
  EmbedParams p;
  p.item_table = item_table; p.pos_table = pos_table; p.aux_table = aux_table; p.seq = seq; p.aux_ids = aux_ids;
  p.T = B * L; p.L = L; p.D = D; p.F = F; p.mode = mode; p.n_rows = n_rows; p.n_aux = n_aux; p.item_scale = item_scale;
  p.ln_w = ln_w; p.ln_b = ln_b; p.eps = eps; p.x0 = (bf16*)x0_bf16; p.x0_f32 = x0_f32; p.q = (bf16*)q_bf16;
  p.stats = stats; p.ldx = ldx;
  p.drop_seed = drop_seed; p.drop_stream = drop_stream; p.drop_step = drop_step;
  p.drop_thresh = drop_p > 0.f ? (uint32_t)((double)drop_p * 4294967296.0) : 0;
  p.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  embed_ln_kernel<<<grid_for_warps(p.T, 8), 256, 0, (cudaStream_t)stream>>>(p);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_layernorm_fwd(const void* x, int ldx, const float* w, const float* b, float eps, void* y_bf16,
                                   float* y_f32, int ldy, float* stats, int64_t T, int H, int64_t row_stride,
                                   int64_t row_offset, void* stream) {
  SRFRD_REQUIRE(x && w && b && (y_bf16 || y_f32), "layernorm_fwd: null pointer");
  SRFRD_REQUIRE(H <= 32 * MAXC, "layernorm_fwd: width %d unsupported", H);
  if (T == 0) return 0;
  LnFwdParams p;
  p.x = (const bf16*)x; p.ldx = ldx; p.w = w; p.b = b; p.eps = eps; p.y_bf16 = (bf16*)y_bf16; p.y_f32 = y_f32;
  p.ldy = ldy; p.stats = stats; p.T = T; p.H = H; p.row_stride = row_stride; p.row_offset = row_offset;
  ln_fwd_kernel<<<grid_for_warps(T, 8), 256, 0, (cudaStream_t)stream>>>(p);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_layernorm_bwd(const void* dy_bf16, const float* dy_f32, int lddy, const void* x, int ldx,
                                   const float* stats, const float* w, const void* add, int ldadd,
                                   const int64_t* row_ids, void* dx, int lddx, float* dw, float* db, int64_t T, int H,
                                   void* stream) {
  SRFRD_REQUIRE((dy_bf16 || dy_f32) && x && stats && w && dx && dw && db, "layernorm_bwd: null pointer");
  SRFRD_REQUIRE(H <= 32 * MAXC, "layernorm_bwd: width %d unsupported", H);
  if (T == 0) return 0;
  LnBwdParams p;
  p.dy_bf16 = (const bf16*)dy_bf16; p.dy_f32 = dy_f32; p.lddy = lddy; p.x = (const bf16*)x; p.ldx = ldx;
  p.stats = stats; p.w = w; p.add = (const bf16*)add; p.ldadd = ldadd; p.row_ids = row_ids; p.dx = (bf16*)dx;
  p.lddx = lddx; p.dw = dw; p.db = db; p.T = T; p.H = H;
  int grid = grid_for_warps(T, 8);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  ln_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
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
  colsum_kernel<<<grid, block, 256 * 2 * sizeof(float), (cudaStream_t)stream>>>((const bf16*)X, M, N, ld, out,
                                                                               (int)rows_per_block);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_cast_weights(const srfrd_cast_desc_t* descs_dev, int n, void* stream) {
  SRFRD_REQUIRE(descs_dev || n == 0, "cast_weights: null table");
  if (n == 0) return 0;
  cast_weights_kernel<<<dim3(n, 8), 256, 0, (cudaStream_t)stream>>>(descs_dev, n);
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
  dropout_apply_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, ldx, (bf16*)out, ldo, M, N, seed, (uint32_t)((double)drop_p * 4294967296.0), stream_id,
      1.f / (1.f - drop_p), drop_step);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

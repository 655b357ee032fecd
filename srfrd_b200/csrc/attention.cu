// Causal self-attention over one (sequence, head) per CTA  --  SURVEY.md 2.3 row k2.
// Restates F.multi_head_attention_forward's need_weights branch as called from SRFR_model.py:112:
//   S = (q * hd^-1/2) k^T, -inf above the diagonal, NO key-padding mask (:113), softmax, dropout, @ v.
// The (L x L) probability matrix lives in shared memory only; q/k/v stream through in feature chunks,
// so backward recomputes P from q,k instead of reading a saved (B, L, L) tensor from HBM.
// Round-1 implementation: fp32 SIMT math on bf16 operands (about 10 % of the step's FLOPs).
#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

static constexpr int CH = 32;        // feature chunk
static constexpr int CHP = CH + 2;   // padded row (bf16) -> 17 words, conflict-free across rows
static constexpr int ATT_THREADS = 256;

struct AttnParams {
  const bf16 *q, *k, *v;
  int ldq, ldkv;
  bf16* o; int ldo;                   // forward output
  const bf16* dout; int lddo;         // backward input
  bf16 *dq, *dk, *dv; int lddq, lddkv;
  int L, hd, heads;
  float scale;
  uint64_t drop_seed; uint32_t drop_thresh, drop_stream; float drop_scale;
  const float* drop_step;
};

__device__ __forceinline__ void load_chunk(bf16* dst, const bf16* src, int ld, int L, int c0, int hd) {
  // dst[L][CHP] <- src[l*ld + c0 + c], zero beyond hd; 2 bf16 per thread-iteration
  for (int i = threadIdx.x; i < L * (CH / 2); i += blockDim.x) {
    const int l = i / (CH / 2), c = (i % (CH / 2)) * 2;
    uint32_t w = 0;
    if (c0 + c < hd) w = *reinterpret_cast<const uint32_t*>(src + (size_t)l * ld + c0 + c);
    *reinterpret_cast<uint32_t*>(dst + l * CHP + c) = w;
  }
}

// The causal (L x L) matrices are stored as packed lower triangles: row i starts at i (i + 1) / 2.
__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

// S[i][j] += sum_c a[i][c] * b[j][c] for j <= i  (strips of 1 x 4)
__device__ __forceinline__ void accum_scores(float* S, const bf16* a, const bf16* b, int L) {
  const int strips = (L + 3) / 4;
  for (int w = threadIdx.x; w < L * strips; w += blockDim.x) {
    const int i = w / strips, j0 = (w % strips) * 4;
    if (j0 > i) continue;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat162* ar = reinterpret_cast<const __nv_bfloat162*>(a + i * CHP);
#pragma unroll 4
    for (int c = 0; c < CH / 2; ++c) {
      const float2 av = __bfloat1622float2(ar[c]);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = min(j0 + jj, L - 1);
        const float2 bv = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(b + j * CHP)[c]);
        acc[jj] = fmaf(av.x, bv.x, fmaf(av.y, bv.y, acc[jj]));
      }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      if (j0 + jj <= i) S[tri(i) + j0 + jj] += acc[jj];
  }
}

__device__ __forceinline__ void zero_f32(float* p, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0.f;
}

// rows of S -> softmax probabilities (pre-dropout), one warp per row
__device__ __forceinline__ void softmax_rows(float* S, int L, float scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = warp; i < L; i += nw) {
    float m = -INFINITY;
    for (int j = lane; j <= i; j += 32) m = fmaxf(m, S[tri(i) + j] * scale);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j <= i; j += 32) {
      const float e = __expf(S[tri(i) + j] * scale - m);
      S[tri(i) + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j <= i; j += 32) S[tri(i) + j] *= inv;
  }
}

__device__ __forceinline__ float drop_factor(const AttnParams& p, int bh, int i, int j) {
  if (!p.drop_thresh) return 1.f;
  const uint64_t idx = ((uint64_t)bh * p.L + i) * p.L + j;
  return dropout_keep(p.drop_seed, p.drop_stream, idx, p.drop_thresh) ? p.drop_scale : 0.f;
}

__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int L = p.L, NTRI = (tri(L) + 3) & ~3;
  float* S = reinterpret_cast<float*>(smem);
  bf16* bufA = reinterpret_cast<bf16*>(S + NTRI);
  bf16* bufB = bufA + L * CHP;
  const int bh = blockIdx.x, b = bh / p.heads, h = bh % p.heads;
  const bf16* q = p.q + (size_t)b * L * p.ldq + h * p.hd;
  const bf16* k = p.k + (size_t)b * L * p.ldkv + h * p.hd;
  const bf16* v = p.v + (size_t)b * L * p.ldkv + h * p.hd;
  bf16* o = p.o + (size_t)b * L * p.ldo + h * p.hd;
  if (p.drop_thresh) p.drop_seed = mix_seed(p.drop_seed, p.drop_step);

  zero_f32(S, NTRI);
  for (int c0 = 0; c0 < p.hd; c0 += CH) {
    __syncthreads();
    load_chunk(bufA, q, p.ldq, L, c0, p.hd);
    load_chunk(bufB, k, p.ldkv, L, c0, p.hd);
    __syncthreads();
    accum_scores(S, bufA, bufB, L);
  }
  __syncthreads();
  softmax_rows(S, L, p.scale);
  __syncthreads();
  if (p.drop_thresh) {
    for (int w = threadIdx.x; w < L * L; w += blockDim.x) {
      const int i = w / L, j = w % L;
      if (j <= i) S[tri(i) + j] *= drop_factor(p, bh, i, j);
    }
  }
  // O[i][c] = sum_{j<=i} P[i][j] v[j][c]
  for (int c0 = 0; c0 < p.hd; c0 += CH) {
    __syncthreads();
    load_chunk(bufA, v, p.ldkv, L, c0, p.hd);
    __syncthreads();
    for (int w = threadIdx.x; w < L * (CH / 2); w += blockDim.x) {
      const int i = w / (CH / 2), c = (w % (CH / 2)) * 2;
      if (c0 + c >= p.hd) continue;
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j <= i; ++j) {
        const float pij = S[tri(i) + j];
        const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(bufA + j * CHP + c));
        a0 = fmaf(pij, vv.x, a0);
        a1 = fmaf(pij, vv.y, a1);
      }
      *reinterpret_cast<uint32_t*>(o + (size_t)i * p.ldo + c0 + c) = pack_bf16x2(a0, a1);
    }
  }
}

// Backward: recompute A = softmax(S); dAd = dO V^T; delta_i = sum_j dAd*Ad; dS = A*(dAd*M - delta)*scale;
//           dV = Ad^T dO; dq = dS k; dk = dS^T q.
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_kernel(AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int L = p.L, NTRI = (tri(L) + 3) & ~3;
  float* A = reinterpret_cast<float*>(smem);        // probabilities, later Ad = A * M
  float* G = A + NTRI;                              // dAd, later dS
  bf16* bufA = reinterpret_cast<bf16*>(G + NTRI);
  bf16* bufB = bufA + L * CHP;
  const int bh = blockIdx.x, b = bh / p.heads, h = bh % p.heads;
  const bf16* q = p.q + (size_t)b * L * p.ldq + h * p.hd;
  const bf16* k = p.k + (size_t)b * L * p.ldkv + h * p.hd;
  const bf16* v = p.v + (size_t)b * L * p.ldkv + h * p.hd;
  const bf16* dout = p.dout + (size_t)b * L * p.lddo + h * p.hd;
  bf16* dq = p.dq + (size_t)b * L * p.lddq + h * p.hd;
  bf16* dk = p.dk + (size_t)b * L * p.lddkv + h * p.hd;
  bf16* dv = p.dv + (size_t)b * L * p.lddkv + h * p.hd;
  if (p.drop_thresh) p.drop_seed = mix_seed(p.drop_seed, p.drop_step);

  zero_f32(A, 2 * NTRI);
  for (int c0 = 0; c0 < p.hd; c0 += CH) {
    __syncthreads();
    load_chunk(bufA, q, p.ldq, L, c0, p.hd);
    load_chunk(bufB, k, p.ldkv, L, c0, p.hd);
    __syncthreads();
    accum_scores(A, bufA, bufB, L);
    __syncthreads();
    load_chunk(bufA, dout, p.lddo, L, c0, p.hd);
    load_chunk(bufB, v, p.ldkv, L, c0, p.hd);
    __syncthreads();
    accum_scores(G, bufA, bufB, L);
  }
  __syncthreads();
  softmax_rows(A, L, p.scale);
  __syncthreads();
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = warp; i < L; i += nw) {
      float delta = 0.f;
      for (int j = lane; j <= i; j += 32) {
        const float m = drop_factor(p, bh, i, j);
        const float gm = G[tri(i) + j] * m;
        G[tri(i) + j] = gm;                    // dA = dAd * M
        delta += gm * A[tri(i) + j];
        // keep A (undropped) until dS is formed; Ad is rebuilt below
      }
      delta = warp_sum(delta);
      for (int j = lane; j <= i; j += 32) {
        const float a = A[tri(i) + j];
        G[tri(i) + j] = a * (G[tri(i) + j] - delta) * p.scale;   // dS (scale folded: S = scale * q k^T)
        A[tri(i) + j] = a * drop_factor(p, bh, i, j);            // Ad
      }
    }
  }
  for (int c0 = 0; c0 < p.hd; c0 += CH) {
    // dV[j][c] = sum_{i>=j} Ad[i][j] dO[i][c];  dK[j][c] = sum_{i>=j} dS[i][j] q[i][c]
    __syncthreads();
    load_chunk(bufA, dout, p.lddo, L, c0, p.hd);
    load_chunk(bufB, q, p.ldq, L, c0, p.hd);
    __syncthreads();
    for (int w = threadIdx.x; w < L * (CH / 2); w += blockDim.x) {
      const int j = w / (CH / 2), c = (w % (CH / 2)) * 2;
      if (c0 + c >= p.hd) continue;
      float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
      for (int i = j; i < L; ++i) {
        const float ad = A[tri(i) + j], ds = G[tri(i) + j];
        const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(bufA + i * CHP + c));
        const float2 qq = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(bufB + i * CHP + c));
        v0 = fmaf(ad, d.x, v0); v1 = fmaf(ad, d.y, v1);
        k0 = fmaf(ds, qq.x, k0); k1 = fmaf(ds, qq.y, k1);
      }
      *reinterpret_cast<uint32_t*>(dv + (size_t)j * p.lddkv + c0 + c) = pack_bf16x2(v0, v1);
      *reinterpret_cast<uint32_t*>(dk + (size_t)j * p.lddkv + c0 + c) = pack_bf16x2(k0, k1);
    }
    // dQ[i][c] = sum_{j<=i} dS[i][j] k[j][c]
    __syncthreads();
    load_chunk(bufA, k, p.ldkv, L, c0, p.hd);
    __syncthreads();
    for (int w = threadIdx.x; w < L * (CH / 2); w += blockDim.x) {
      const int i = w / (CH / 2), c = (w % (CH / 2)) * 2;
      if (c0 + c >= p.hd) continue;
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j <= i; ++j) {
        const float ds = G[tri(i) + j];
        const float2 kk = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(bufA + j * CHP + c));
        a0 = fmaf(ds, kk.x, a0); a1 = fmaf(ds, kk.y, a1);
      }
      *reinterpret_cast<uint32_t*>(dq + (size_t)i * p.lddq + c0 + c) = pack_bf16x2(a0, a1);
    }
  }
}

static int host_ntri(int L) { return ((L * (L + 1) / 2) + 3) & ~3; }

static int fill_common(AttnParams& p, int L, int H, int heads, float drop_p, uint64_t seed, uint32_t stream_id,
                       const float* drop_step) {
  p.drop_step = drop_step;
  SRFRD_REQUIRE(heads > 0 && H % heads == 0, "attention: hidden %d not divisible by heads %d", H, heads);
  p.L = L; p.heads = heads; p.hd = H / heads;
  SRFRD_REQUIRE(p.hd % 2 == 0, "attention: head_dim must be even");
  p.scale = 1.0f / sqrtf((float)p.hd);
  p.drop_seed = seed; p.drop_stream = stream_id; p.drop_thresh = 0; p.drop_scale = 1.f;
  if (drop_p > 0.f) {
    SRFRD_REQUIRE(drop_p < 1.f, "attention: dropout p must be < 1");
    p.drop_thresh = (uint32_t)((double)drop_p * 4294967296.0);
    p.drop_scale = 1.f / (1.f - drop_p);
  }
  return 0;
}

// SIMT path: shapes the tcgen05 kernels (attention_tc.cu) do not cover -- maxlen > 128 or head_dim % 16 != 0.
// Packed-triangle score storage keeps maxlen 200 (C4) inside 227 KB: forward up to L ~ 305, backward up to L = 223.
int attn_fwd_simt(const void* q, int ldq, const void* k, const void* v, int ldkv, void* o, int ldo, int64_t B, int L,
                  int H, int heads, float drop_p, uint64_t seed, uint32_t stream_id, const float* drop_step,
                  void* stream) {
  SRFRD_REQUIRE(q && k && v && o, "attention_fwd: null pointer");
  SRFRD_REQUIRE(ldq % 2 == 0 && ldkv % 2 == 0 && ldo % 2 == 0, "attention_fwd: leading dims must be even");
  if (B == 0 || L == 0) return 0;
  AttnParams p = {};
  if (int rc = fill_common(p, L, H, heads, drop_p, seed, stream_id, drop_step)) return rc;
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.ldq = ldq; p.ldkv = ldkv;
  p.o = (bf16*)o; p.ldo = ldo;
  const size_t smem = (size_t)host_ntri(L) * 4 + 2 * (size_t)L * CHP * 2;
  SRFRD_REQUIRE(smem <= 227 * 1024, "attention_fwd: maxlen %d needs %zu B of shared memory (> 227 KB)", L, smem);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    SRFRD_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  attn_fwd_kernel<<<(unsigned)(B * heads), ATT_THREADS, smem, (cudaStream_t)stream>>>(p);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

int attn_bwd_simt(const void* dout, int lddo, const void* q, int ldq, const void* k, const void* v, int ldkv, void* dq,
                  int lddq, void* dk, void* dv, int lddkv, int64_t B, int L, int H, int heads, float drop_p,
                  uint64_t seed, uint32_t stream_id, const float* drop_step, void* stream) {
  SRFRD_REQUIRE(dout && q && k && v && dq && dk && dv, "attention_bwd: null pointer");
  SRFRD_REQUIRE(lddo % 2 == 0 && ldq % 2 == 0 && ldkv % 2 == 0 && lddq % 2 == 0 && lddkv % 2 == 0,
                "attention_bwd: leading dims must be even");
  if (B == 0 || L == 0) return 0;
  AttnParams p = {};
  if (int rc = fill_common(p, L, H, heads, drop_p, seed, stream_id, drop_step)) return rc;
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.ldq = ldq; p.ldkv = ldkv;
  p.dout = (const bf16*)dout; p.lddo = lddo; p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv;
  p.lddq = lddq; p.lddkv = lddkv;
  const size_t smem = 2 * (size_t)host_ntri(L) * 4 + 2 * (size_t)L * CHP * 2;
  SRFRD_REQUIRE(smem <= 227 * 1024, "attention_bwd: maxlen %d needs %zu B of shared memory (> 227 KB)", L, smem);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    SRFRD_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  attn_bwd_kernel<<<(unsigned)(B * heads), ATT_THREADS, smem, (cudaStream_t)stream>>>(p);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

}  // namespace srfrd

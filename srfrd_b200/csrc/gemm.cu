// tcgen05 / TMEM / TMA GEMMs for the encoder's dense contractions.
//
//  gemm_tn   : C[M,N] = epilogue(A[M,K] * B[N,K]^T)   both operands K-major (row-major, K contiguous)
//              forward linears (A = activations, B = weight (out,in)) and data-gradients
//              (A = dY, B = W^T shadow).  Persistent CTAs, 128-row tiles, TMEM double buffering.
//  gemm_wgrad: dW[Mo,No] += sum_t A[t,Mo] * B[t,No]   both operands MN-major (token index = K)
//              weight gradients, split-K over tokens, fp32 red.add into the flat gradient buffer.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_idx & 3).
#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;                       // 64 bf16 = one 128-byte swizzle row
static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
static constexpr int GEMM_THREADS = 192;
static constexpr int TMEM_COLS = 512;

struct GemmEpilogue {
  const float* bias;        // [N] or null
  const bf16* residual;     // [M, ldr] or null
  const bf16* gate;         // [M, ldg] or null : v *= (gate > 0)      (ReLU backward)
  const int64_t* row_ids;   // [M] or null      : v *= (row_ids[m] != 0) (pad re-mask, SRFR_model.py:121)
  bf16* out_bf16;           // [M, ldc] or null
  float* out_f32;           // [M, ldc] or null
  int ldr, ldg, ldc;
  int relu;
  uint64_t drop_seed;       // dropout (drop_thresh == 0 -> off): v = keep ? v * drop_scale : 0
  uint32_t drop_thresh, drop_stream;
  float drop_scale;
  const float* drop_step;
};

struct GemmShape {
  int M, N, K;
  int block_n, m_tiles, n_tiles, stages;
};

__device__ __forceinline__ float apply_epilogue(float v, int j, const GemmEpilogue& e, const float* bias_v,
                                                const float* res_v, const float* gate_v, float rowm,
                                                uint64_t elem_idx) {
  if (e.bias) v += bias_v[j];
  if (e.drop_thresh) v = dropout_keep(e.drop_seed, e.drop_stream, elem_idx, e.drop_thresh) ? v * e.drop_scale : 0.f;
  if (e.relu) v = fmaxf(v, 0.f);
  if (e.gate) v = gate_v[j] > 0.f ? v : 0.f;
  if (e.residual) v += res_v[j];
  return v * rowm;
}

__device__ __forceinline__ void load16_bf16(const bf16* p, float* out) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    out[2 * i] = __low2float(t);
    out[2 * i + 1] = __high2float(t);
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmShape s,
               GemmEpilogue e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage_bytes = s.block_n * BLOCK_K * 2;
  uint8_t* smA = smem;
  uint8_t* smB = smem + s.stages * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + s.stages * b_stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + s.stages;
  uint64_t* tfull = bars + 2 * s.stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = s.m_tiles * s.n_tiles;
  const int kblocks = (s.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < s.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Warps 0 and 1 stay converged; only the issuing instructions are predicated on elect.sync so that barrier
  // addresses and descriptors live in uniform registers (see topk.cu for the measurement behind this).
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m0 = (t / s.n_tiles) * BLOCK_M, n0 = (t % s.n_tiles) * s.block_n;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], A_STAGE_BYTES + b_stage_bytes);
          tma_load_2d(smA + stage * A_STAGE_BYTES, &tmA, &full[stage], kb * BLOCK_K, m0, SRFRD_EVICT_FIRST);
          tma_load_2d(smB + stage * b_stage_bytes, &tmB, &full[stage], kb * BLOCK_K, n0, SRFRD_EVICT_LAST);
        }
        __syncwarp();
        if (++stage == s.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    const uint64_t adesc0 = umma_smem_desc(smem_u32(smA), 0, 1024), bdesc0 = umma_smem_desc(smem_u32(smB), 0, 1024);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int n0 = (t % s.n_tiles) * s.block_n;
      int bn = min(s.block_n, s.N - n0);
      bn = (bn + 15) & ~15;
      const uint32_t idesc = umma_idesc_bf16(BLOCK_M, bn, 0, 0);
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * 256;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        // K-major SW128: 8-row groups are 1024 B apart (SBO); +32 B (= 2 in descriptor units) per UMMA_K = 16
        const uint64_t ad = adesc0 + (uint64_t)(stage * (A_STAGE_BYTES >> 4));
        const uint64_t bd = bdesc0 + (uint64_t)(stage * (b_stage_bytes >> 4));
        const int ksteps = min(BLOCK_K / 16, (s.K - kb * BLOCK_K + 15) / 16);
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k) umma_bf16(tacc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
          if (kb == kblocks - 1) umma_commit(&tfull[as]);
        }
        __syncwarp();
        if (++stage == s.stages) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    const int quarter = warp & 3;
    int as = 0; uint32_t aphase = 0;
    if (e.drop_thresh) e.drop_seed = mix_seed(e.drop_seed, e.drop_step);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m0 = (t / s.n_tiles) * BLOCK_M, n0 = (t % s.n_tiles) * s.block_n;
      const int bn = min(s.block_n, s.N - n0);
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < s.M;
      float rowm = 1.f;
      if (e.row_ids && row_ok) rowm = (__ldg(e.row_ids + row) != 0) ? 1.f : 0.f;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * 256;
      for (int c = 0; c < bn; c += 16) {
        uint32_t raw[16];
        tmem_ld16(taddr + c, raw);
        tmem_ld_wait();
        if (row_ok) {
          const int n = n0 + c;
          float bias_v[16], res_v[16], gate_v[16], v[16];
          if (e.bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + n + j));
              bias_v[j] = b4.x; bias_v[j + 1] = b4.y; bias_v[j + 2] = b4.z; bias_v[j + 3] = b4.w;
            }
          }
          if (e.residual) load16_bf16(e.residual + (size_t)row * e.ldr + n, res_v);
          if (e.gate) load16_bf16(e.gate + (size_t)row * e.ldg + n, gate_v);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = apply_epilogue(__uint_as_float(raw[j]), j, e, bias_v, res_v, gate_v, rowm,
                                  (uint64_t)row * (uint64_t)s.N + (uint64_t)(n + j));
          if (e.out_bf16) {
            uint4* o = reinterpret_cast<uint4*>(e.out_bf16 + (size_t)row * e.ldc + n);
            o[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                              pack_bf16x2(v[6], v[7]));
            o[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                              pack_bf16x2(v[14], v[15]));
          }
          if (e.out_f32) {
            float4* o = reinterpret_cast<float4*>(e.out_f32 + (size_t)row * e.ldc + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: split-K over tokens, MN-major operands
// ---------------------------------------------------------------------------------------------
struct WgradShape {
  int T, Mo, No;       // tokens, output rows (dY width), output cols (X width)
  int block_n;         // columns per CTA tile (multiple of 16, <= 256)
  int a_atoms, b_atoms;  // 64-wide MN atoms per stage
  int stages, k_splits;
  float* out;          // [Mo, ldw] fp32, accumulated with red.add
  int ldw;
  float* dbias;        // [Mo] or null: column sums of dY, computed by one extra N=16 MMA per K step against a
                       // constant all-ones operand (the bias gradient for free, no second pass over dY)
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, WgradShape s) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int ATOM_BYTES = BLOCK_K * 128;             // 64 k-rows x 128 B (64 bf16 along MN)
  const int a_bytes = s.a_atoms * ATOM_BYTES, b_bytes = s.b_atoms * ATOM_BYTES;
  uint8_t* smA = smem;
  uint8_t* smB = smem + s.stages * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + s.stages * b_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + s.stages;
  uint64_t* tfull = bars + 2 * s.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  uint8_t* ones = smem + s.stages * (a_bytes + b_bytes) + 1024;      // 64 k-rows x 128 B of bf16 1.0 (after the barriers)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.z * s.block_n;
  const bool do_bias = s.dbias != nullptr && blockIdx.z == 0;
  if (do_bias) {
    for (int i = threadIdx.x; i < ATOM_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
    fence_proxy_async();
  }
  const int kblocks_total = (s.T + BLOCK_K - 1) / BLOCK_K;
  const int per = (kblocks_total + s.k_splits - 1) / s.k_splits;
  const int kb_begin = blockIdx.x * per, kb_end = min(kblocks_total, kb_begin + per);
  const int nkb = max(0, kb_end - kb_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < s.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nkb > 0) {
    if (warp == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], a_bytes + b_bytes);
          for (int a = 0; a < s.a_atoms; ++a)
            tma_load_2d(smA + stage * a_bytes + a * ATOM_BYTES, &tmA, &full[stage], m0 + a * 64, kb * BLOCK_K,
                        SRFRD_EVICT_FIRST);
          for (int b = 0; b < s.b_atoms; ++b)
            tma_load_2d(smB + stage * b_bytes + b * ATOM_BYTES, &tmB, &full[stage], n0 + b * 64, kb * BLOCK_K,
                        SRFRD_EVICT_FIRST);
        }
        __syncwarp();
        if (++stage == s.stages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1) {
      int bn = min(s.block_n, s.No - n0);
      bn = (bn + 15) & ~15;
      const uint32_t idesc = umma_idesc_bf16(BLOCK_M, bn, 1, 1);
      // MN-major SW128: 64-wide MN atoms are ATOM_BYTES apart (LBO); 8 k-rows = 1024 B (SBO);
      // one UMMA_K = 16 k-rows = 2048 B (= 128 in descriptor units)
      const uint64_t adesc0 = umma_smem_desc(smem_u32(smA), ATOM_BYTES, 1024);
      const uint64_t bdesc0 = umma_smem_desc(smem_u32(smB), ATOM_BYTES, 1024);
      const uint64_t onesdesc = umma_smem_desc(smem_u32(ones), ATOM_BYTES, 1024);
      const uint32_t idesc_ones = umma_idesc_bf16(BLOCK_M, 16, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t ad = adesc0 + (uint64_t)(stage * (a_bytes >> 4)), bd = bdesc0 + (uint64_t)(stage * (b_bytes >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_bf16(tmem_base, ad + 128 * k, bd + 128 * k, idesc, (kb > kb_begin) || (k > 0));
          if (do_bias) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k)
              umma_bf16(tmem_base + 240, ad + 128 * k, onesdesc + 128 * k, idesc_ones, (kb > kb_begin) || (k > 0));
          }
          umma_commit(&empty[stage]);
          if (kb == kb_end - 1) umma_commit(tfull);
        }
        __syncwarp();
        if (++stage == s.stages) { stage = 0; phase ^= 1; }
      }
    } else {
      const int quarter = warp & 3;
      const int row = m0 + quarter * 32 + lane;
      const int bn = min(s.block_n, s.No - n0);
      mbar_wait(tfull, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      for (int c = 0; c < bn; c += 16) {
        uint32_t raw[16];
        tmem_ld16(taddr + c, raw);
        tmem_ld_wait();
        if (row < s.Mo) {
          float* o = s.out + (size_t)row * s.ldw + n0 + c;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c + j < s.No) red_add_f32(o + j, __uint_as_float(raw[j]));
        }
      }
      if (do_bias) {                       // columns [240, 256) all hold sum_t dY[t, row]
        uint32_t raw[16];
        tmem_ld16(taddr + 240, raw);
        tmem_ld_wait();
        if (row < s.Mo) red_add_f32(s.dbias + row, __uint_as_float(raw[0]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ---------------------------------------------------------------------------------------------
// plain SIMT reference GEMM: used ONLY by tests to cross-check the tensor-core kernels on device
// ---------------------------------------------------------------------------------------------
__global__ void gemm_ref_kernel(const bf16* A, int lda, const bf16* B, int ldb, float* C, int ldc, int M, int N, int K,
                                int a_mn_major, int b_mn_major) {
  int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float a = a_mn_major ? bf2f(A[(size_t)k * lda + m]) : bf2f(A[(size_t)m * lda + k]);
    float b = b_mn_major ? bf2f(B[(size_t)k * ldb + n]) : bf2f(B[(size_t)n * ldb + k]);
    acc = fmaf(a, b, acc);
  }
  C[(size_t)m * ldc + n] = acc;
}

static int pick_block_n(int N) {
  int tiles = (N + 255) / 256;
  int bn = ((N + tiles - 1) / tiles + 15) & ~15;
  return bn;
}

}  // namespace srfrd

using namespace srfrd;

extern "C" int srfrd_gemm_tn(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                             const srfrd_gemm_epilogue_t* ep, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRFRD_REQUIRE(A && B && ep, "gemm_tn: null operand");
  SRFRD_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_tn: empty shape M=%d N=%d K=%d", M, N, K);
  SRFRD_REQUIRE(N % 16 == 0 && K % 8 == 0, "gemm_tn: need N %% 16 == 0 and K %% 8 == 0 (got N=%d K=%d)", N, K);
  SRFRD_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= K, "gemm_tn: bad leading dims lda=%d ldb=%d", lda, ldb);
  SRFRD_REQUIRE(ep->out_bf16 || ep->out_f32, "gemm_tn: no output");
  SRFRD_REQUIRE(ep->ldc % 8 == 0 && ep->ldc >= N, "gemm_tn: bad ldc=%d", ep->ldc);
  SRFRD_REQUIRE(!ep->residual || ep->ldr % 8 == 0, "gemm_tn: bad ldr");
  SRFRD_REQUIRE(!ep->gate || ep->ldg % 8 == 0, "gemm_tn: bad ldg");
  GemmShape s;
  s.M = M; s.N = N; s.K = K;
  s.block_n = pick_block_n(N);
  s.n_tiles = (N + s.block_n - 1) / s.block_n;
  s.m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int stage_bytes = A_STAGE_BYTES + s.block_n * BLOCK_K * 2;
  s.stages = (200 * 1024) / stage_bytes;
  if (s.stages > 6) s.stages = 6;
  const size_t smem = (size_t)s.stages * stage_bytes + 1024 + 256;
  CUtensorMap tmA, tmB;
  if (int rc = make_tmap_bf16_2d(&tmA, A, M, K, lda, BLOCK_M, BLOCK_K)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmB, B, N, K, ldb, s.block_n, BLOCK_K)) return rc;
  GemmEpilogue e;
  e.bias = ep->bias; e.residual = (const bf16*)ep->residual; e.gate = (const bf16*)ep->gate;
  e.row_ids = ep->row_ids; e.out_bf16 = (bf16*)ep->out_bf16; e.out_f32 = ep->out_f32;
  e.ldr = ep->ldr; e.ldg = ep->ldg; e.ldc = ep->ldc; e.relu = ep->relu;
  e.drop_seed = ep->drop_seed; e.drop_stream = ep->drop_stream; e.drop_step = ep->drop_step;
  e.drop_thresh = 0; e.drop_scale = 1.f;
  if (ep->drop_p > 0.f) {
    SRFRD_REQUIRE(ep->drop_p < 1.f, "gemm_tn: dropout p must be < 1");
    e.drop_thresh = (uint32_t)((double)ep->drop_p * 4294967296.0);
    e.drop_scale = 1.f / (1.f - ep->drop_p);
  }
  static bool attr_set = false;
  if (!attr_set) {
    SRFRD_CUDA(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int grid = s.m_tiles * s.n_tiles;
  if (grid > num_sms()) grid = num_sms();
  gemm_tn_kernel<<<grid, GEMM_THREADS, smem, stream>>>(tmA, tmB, s, e);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_gemm_wgrad(const void* dY, int lda, const void* X, int ldb, int64_t T, int Mo, int No,
                                float* dW, int ldw, float* dbias, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SRFRD_REQUIRE(dY && X && dW, "gemm_wgrad: null operand");
  SRFRD_REQUIRE(T > 0 && Mo > 0 && No > 0, "gemm_wgrad: empty shape");
  SRFRD_REQUIRE(Mo % 8 == 0 && No % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "gemm_wgrad: widths must be multiples of 8");
  SRFRD_REQUIRE(T < (1ll << 31), "gemm_wgrad: too many tokens");
  WgradShape s;
  s.T = (int)T; s.Mo = Mo; s.No = No;
  s.block_n = pick_block_n((No + 15) & ~15);
  const int n_tiles = (No + s.block_n - 1) / s.block_n, m_tiles = (Mo + BLOCK_M - 1) / BLOCK_M;
  s.a_atoms = 2;
  s.b_atoms = (s.block_n + 63) / 64;
  const int stage_bytes = (s.a_atoms + s.b_atoms) * BLOCK_K * 128;
  s.stages = (190 * 1024) / stage_bytes;
  if (s.stages > 6) s.stages = 6;
  const int kblocks = (s.T + BLOCK_K - 1) / BLOCK_K;
  int splits = num_sms() / (n_tiles * m_tiles);
  if (splits < 1) splits = 1;
  if (splits > kblocks) splits = kblocks;
  s.k_splits = splits;
  s.out = dW; s.ldw = ldw; s.dbias = dbias;
  SRFRD_REQUIRE(!dbias || s.block_n <= 240, "gemm_wgrad: fused bias gradient needs a column tile <= 240");
  const size_t smem = (size_t)s.stages * stage_bytes + 1024 + 1024 + BLOCK_K * 128;
  CUtensorMap tmA, tmB;
  // MN-major: the TMA box is [64 tokens (rows), 64 features (cols)]
  if (int rc = make_tmap_bf16_2d(&tmA, dY, T, Mo, lda, BLOCK_K, 64)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmB, X, T, No, ldb, BLOCK_K, 64)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    SRFRD_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid(s.k_splits, m_tiles, n_tiles);
  gemm_wgrad_kernel<<<grid, GEMM_THREADS, smem, stream>>>(tmA, tmB, s);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

extern "C" int srfrd_gemm_ref(const void* A, int lda, const void* B, int ldb, float* C, int ldc, int M, int N, int K,
                              int a_mn_major, int b_mn_major, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  dim3 grid((N + 127) / 128, M);
  gemm_ref_kernel<<<grid, 128, 0, stream>>>((const bf16*)A, lda, (const bf16*)B, ldb, C, ldc, M, N, K, a_mn_major,
                                            b_mn_major);
  SRFRD_LAUNCH_CHECK();
  return 0;
}

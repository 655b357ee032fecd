// Ceiling for K1: how fast can this HBM serve random 256-byte rows (ids streamed from HBM) with an equal sequential
// write stream, with nothing else in the kernel?  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_peak gather_peak.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <random>

template <int UNROLL>
__global__ void __launch_bounds__(256) gather_rows(const int64_t* __restrict__ ids, const float4* __restrict__ table,
                                                   float4* __restrict__ out, int64_t T) {
  // 16 lanes x 16 B per row; each half-warp handles UNROLL rows per iteration
  const int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4, nhw = ((int64_t)gridDim.x * blockDim.x) >> 4;
  const int sub = threadIdx.x & 15;
  for (int64_t t0 = hw * UNROLL; t0 < T; t0 += nhw * UNROLL) {
    int64_t id[UNROLL];
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) id[u] = t0 + u < T ? __ldg(ids + t0 + u) : 0;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(table + id[u] * 16 + sub);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      if (t0 + u < T) __stcs(out + (t0 + u) * 16 + sub, v[u]);
  }
}

template <int UNROLL>
static float run(const int64_t* ids, const float4* table, float4* out, int64_t T, int blocks_per_sm) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * blocks_per_sm;
  float best = 1e9f;
  for (int it = 0; it < 6; ++it) {
    cudaEventRecord(a);
    gather_rows<UNROLL><<<grid, 256>>>(ids, table, out, T);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (it && ms < best) best = ms;
  }
  return best;
}

int main() {
  const int64_t N = 1000000, T = 81920ll * 50;
  std::vector<int64_t> h(T);
  std::mt19937_64 g(1);
  for (auto& x : h) x = 1 + g() % N;
  int64_t* ids; float4 *table, *out;
  cudaMalloc(&ids, T * 8); cudaMalloc(&table, (N + 1) * 256); cudaMalloc(&out, T * 256);
  cudaMemcpy(ids, h.data(), T * 8, cudaMemcpyHostToDevice);
  cudaMemset(table, 0, (N + 1) * 256);
  const double bytes = (double)T * (8 + 256 + 256);
  for (int bps : {4, 8}) {
    printf("blocks/SM %d: unroll1 %.0f GB/s  unroll2 %.0f GB/s  unroll4 %.0f GB/s  unroll8 %.0f GB/s\n", bps,
           bytes / run<1>(ids, table, out, T, bps) * 1e-6, bytes / run<2>(ids, table, out, T, bps) * 1e-6,
           bytes / run<4>(ids, table, out, T, bps) * 1e-6, bytes / run<8>(ids, table, out, T, bps) * 1e-6);
  }
  // sequential copy of the same volume for comparison
  std::vector<int64_t> seq(T);
  for (int64_t i = 0; i < T; ++i) seq[i] = 1 + i % N;
  cudaMemcpy(ids, seq.data(), T * 8, cudaMemcpyHostToDevice);
  printf("sequential ids, 8 blocks/SM, unroll4: %.0f GB/s\n", bytes / run<4>(ids, table, out, T, 8) * 1e-6);
  return cudaDeviceSynchronize() != cudaSuccess;
}

// Microbenchmark: tcgen05.ld throughput per SM (bytes of registers filled per clock) for several shapes,
// warp counts and the .pack::16b form.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read tmem_read.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_ASM(SHAPE, NREG_STR)                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned." SHAPE ".b32 " NREG_STR ", [%32];"                                        \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),    \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),           \
                 "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),         \
                 "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),         \
                 "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
               : "r"(taddr)                                                                                         \
               : "memory")

#define REGS32                                                                                                      \
  "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29," \
  "%30,%31}"

// MODE 0: 32x32b.x32 (32 cols -> 32 regs)   MODE 1: 32x32b.x32.pack::16b (64 cols -> 32 regs)
// MODE 2: 16x256b.x4? not used.            MODE 3: 32x32b.x16 twice (16 regs each)
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_rd(int iters, int lds_per_wait, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int l = 0; l < lds_per_wait; ++l) {
      uint32_t v[32];
      const uint32_t taddr = base + ((it * lds_per_wait + l) * 64 & 255) + (warp >> 2) * 0;
      if (MODE == 0) LD_ASM("32x32b.x32", REGS32);
      if (MODE == 1) LD_ASM("32x32b.x32.pack::16b", REGS32);
      acc ^= v[0] ^ v[31];
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
  long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_cyc, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 4);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16})
      for (int lpw : {1, 2, 4})
        for (int grid : {1, 148}) {
          for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) tmem_rd<0><<<grid, warps * 32>>>(iters, lpw, d_cyc, d_sink);
            else tmem_rd<1><<<grid, warps * 32>>>(iters, lpw, d_cyc, d_sink);
          }
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148];
          cudaMemcpy(h, d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          const double regs_bytes = (double)warps * iters * lpw * 32 * 32 * 4;   // bytes landed in registers per CTA
          const double cols = (mode == 1 ? 2.0 : 1.0);
          printf("mode=%s warps=%2d lds/wait=%d grid=%3d : %8lld cyc  %.1f reg-B/clk/SM  %.1f tmem-col-B/clk/SM\n",
                 mode ? "x32.pack16" : "x32       ", warps, lpw, grid, mx, regs_bytes / mx, regs_bytes * cols / mx);
        }
  return 0;
}

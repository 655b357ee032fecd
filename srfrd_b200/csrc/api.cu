// C-ABI plumbing: error reporting, device check, TMA descriptor construction.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "srfrd_b200.h"

namespace srfrd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local const int* g_row_limit = nullptr;
const int* row_limit() { return g_row_limit; }

bool pdl_enabled() {
  static int v = -1;
  // Programmatic dependent launch: a kernel's prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps the
  // tail of its predecessor.  Round 1 (dense layout, 0.2 GB per kernel): no gain, 1.83 vs 1.81 ms / step -- the persistent
  // kernels filled every SM.  With the packed token layout a kernel moves ~5 MB and the step is a chain of ~40 launch
  // latencies: 0.558 -> 0.517 ms / step at C2.  On unless SRFRD_PDL=0.
  if (v < 0) { const char* e = getenv("SRFRD_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                      uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  SRFRD_REQUIRE(fn, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  SRFRD_REQUIRE(((uintptr_t)base & 15) == 0, "TMA: base pointer %p is not 16-byte aligned", base);
  SRFRD_REQUIRE((ld * 2) % 16 == 0, "TMA: row pitch %llu elements is not a multiple of 16 bytes", (unsigned long long)ld);
  SRFRD_REQUIRE(box_cols * 2 <= 128 && box_rows <= 256, "TMA: box %u x %u too large", box_rows, box_cols);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRFRD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  return 0;
}

}  // namespace srfrd

extern "C" const char* srfrd_last_error(void) { return srfrd::g_err; }

extern "C" int srfrd_abi_version(void) { return SRFRD_ABI_VERSION; }

extern "C" int srfrd_set_row_limit(const int* rows_dev) {
  srfrd::g_row_limit = rows_dev;
  return 0;
}

extern "C" int srfrd_device_check(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    srfrd::set_error("no CUDA device visible (%s); srfrd_b200 has no CPU fallback", cudaGetErrorString(e));
    return 1;
  }
  int dev = 0, major = 0, minor = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    srfrd::set_error("device compute capability %d.%d is not sm_100 (B200); kernels are built for sm_100a only", major, minor);
    return 2;
  }
  return 0;
}

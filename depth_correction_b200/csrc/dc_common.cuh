// Shared definitions for the depth_correction_b200 CUDA library (sm_100a only).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define DC_HD __host__ __device__ __forceinline__
#define DC_D __device__ __forceinline__
#else
#define DC_HD inline
#define DC_D inline
#endif

#include "../../include/dc_b200.h"

#define DC_WARP 32
#define DC_SLICE 32   // rows per sliced-ELL slice == one warp of query points

#if defined(__CUDACC__)
#define DC_CUDA_CHECK(expr)                          \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) return dc_set_cuda_error(_e, __FILE__, __LINE__); \
  } while (0)

#define DC_LAUNCH_CHECK() DC_CUDA_CHECK(cudaGetLastError())

int dc_set_cuda_error(cudaError_t e, const char* file, int line);
int dc_set_error(int code, const char* msg);

static inline int dc_blocks(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// Sorted map point: xyz in fp64 (exact up-cast of the caller's fp32/fp64 values) and the
// original row index in the bits of w (so tie-breaks and exports need no extra gather).
struct __align__(32) dc_point {
  double x, y, z;
  long long tag;   // original index (search) / unused (step scratch)
};

// 256-bit loads exist on sm_100 (ld.global.v4.f64); one instruction per neighbour gather.
__device__ __forceinline__ dc_point dc_ld_point(const dc_point* p) {
  dc_point r;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=l"(r.tag)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void dc_ld256(const void* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
#endif

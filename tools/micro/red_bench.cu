// Micro-benchmark: throughput of fire-and-forget global reductions by operand type, with the access pattern of
// the scatter-form backward (every thread adds 3 values to 32 neighbours that lie within +-2000 records).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu && ./red_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned int lcg(unsigned int& s) { s = s * 1664525u + 1013904223u; return s; }

template <int MODE>
__global__ void k(double* g64, float* g32, unsigned long long* gi, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned int s = (unsigned int)i * 2654435761u + 12345u;
  for (int c = 0; c < 32; ++c) {
    long long j = i + (long long)(lcg(s) % 4001u) - 2000;
    j = j < 0 ? 0 : (j >= n ? n - 1 : j);
    if (MODE == 0) { atomicAdd(g64 + 4 * j, 1.0); atomicAdd(g64 + 4 * j + 1, 2.0); atomicAdd(g64 + 4 * j + 2, 3.0); }
    if (MODE == 1) { atomicAdd(gi + 4 * j, 1ull); atomicAdd(gi + 4 * j + 1, 2ull); atomicAdd(gi + 4 * j + 2, 3ull); }
    if (MODE == 2) { atomicAdd(g32 + 4 * j, 1.f); atomicAdd(g32 + 4 * j + 1, 2.f); atomicAdd(g32 + 4 * j + 2, 3.f); }
    if (MODE == 3) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g32 + 4 * j), "f"(1.f), "f"(2.f), "f"(3.f), "f"(0.f) : "memory");
    }
    if (MODE == 4) { atomicAdd(g64 + 3 * j, 1.0); atomicAdd(g64 + 3 * j + 1, 2.0); atomicAdd(g64 + 3 * j + 2, 3.0); }
    if (MODE == 5) { atomicAdd(g64 + 4 * j, 1.0); }
  }
}

int main() {
  const long long n = 8366086;
  double* g64; float* g32; unsigned long long* gi;
  cudaMalloc(&g64, n * 32); cudaMalloc(&g32, n * 16); cudaMalloc(&gi, n * 32);
  cudaMemset(g64, 0, n * 32); cudaMemset(g32, 0, n * 16); cudaMemset(gi, 0, n * 32);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[6] = {"3 x red.f64 (32 B stride)", "3 x red.u64 (32 B stride)", "3 x red.f32 (16 B stride)", "1 x red.v4.f32", "3 x red.f64 (24 B stride)", "1 x red.f64"};
  for (int mode = 0; mode < 6; ++mode) {
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      const int blocks = (int)((n + 127) / 128);
      switch (mode) {
        case 0: k<0><<<blocks, 128>>>(g64, g32, gi, n); break;
        case 1: k<1><<<blocks, 128>>>(g64, g32, gi, n); break;
        case 2: k<2><<<blocks, 128>>>(g64, g32, gi, n); break;
        case 3: k<3><<<blocks, 128>>>(g64, g32, gi, n); break;
        case 4: k<4><<<blocks, 128>>>(g64, g32, gi, n); break;
        case 5: k<5><<<blocks, 128>>>(g64, g32, gi, n); break;
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("%-28s %7.3f ms  (%6.1f G edges/s)  %s\n", names[mode], best, n * 32 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

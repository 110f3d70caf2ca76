// Random 2-byte gathers over a 512 MiB table (the scorer's access pattern): which load
// flavour costs the least DRAM traffic / time per miss on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/build/diag_gather scripts/diag_gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int V>
__device__ __forceinline__ uint32_t ld16(const uint16_t* p) {
  uint16_t v;
  if (V == 0) asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
  else if (V == 1) asm volatile("ld.global.u16 %0, [%1];" : "=h"(v) : "l"(p));
  else if (V == 2) asm volatile("ld.global.nc.L2::64B.u16 %0, [%1];" : "=h"(v) : "l"(p));
  else if (V == 3) asm volatile("ld.global.nc.L2::128B.u16 %0, [%1];" : "=h"(v) : "l"(p));
  else if (V == 4) asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
  else if (V == 5) asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  else if (V == 6) asm volatile("ld.global.cs.u16 %0, [%1];" : "=h"(v) : "l"(p));
  else if (V == 7) asm volatile("ld.global.nc.L1::evict_first.u16 %0, [%1];" : "=h"(v) : "l"(p));
  else asm volatile("ld.global.lu.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ uint32_t hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int V, int PER>
__global__ void gather(const uint16_t* __restrict__ t, uint32_t mask, uint32_t* out, uint32_t seed) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t idx[PER], acc = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) idx[k] = hash(g * PER + k + seed) & mask;
#pragma unroll
  for (int k = 0; k < PER; ++k) acc += ld16<V>(t + idx[k]);
  out[g] = acc;
}

template <int V>
static void run(const char* name, const uint16_t* t, uint32_t mask, uint32_t* out, size_t n_gathers) {
  constexpr int PER = 16;
  const int threads = 256;
  const int grid = (int)(n_gathers / PER / threads);
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) gather<V, PER><<<grid, threads>>>(t, mask, out, i * 977u);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  const int reps = 20;
  for (int i = 0; i < reps; ++i) gather<V, PER><<<grid, threads>>>(t, mask, out, 1000u + i * 7919u);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= reps;
  printf("%-46s %8.4f ms  %7.2f G gathers/s  (x32 B = %7.1f GB/s)\n", name, ms,
         n_gathers / (ms * 1e-3) / 1e9, n_gathers * 32.0 / (ms * 1e-3) / 1e9);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const size_t table_bytes = 512ull << 20;
  uint16_t* t; uint32_t* out;
  const size_t n = 1ull << 24;   // gathers per launch
  CK(cudaMalloc(&t, table_bytes)); CK(cudaMemset(t, 1, table_bytes));
  CK(cudaMalloc(&out, n / 16 * 4));
  const uint32_t mask = (uint32_t)(table_bytes / 2 - 1);
  for (size_t gran : {(size_t)0, (size_t)32, (size_t)128}) {
    if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
    size_t v; CK(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
    printf("# cudaLimitMaxL2FetchGranularity = %zu; 2^24 random u16 gathers over 512 MiB, 16 per thread\n", v);
    run<0>("ld.global.nc", t, mask, out, n);
    run<1>("ld.global", t, mask, out, n);
    run<2>("ld.global.nc.L2::64B", t, mask, out, n);
    run<3>("ld.global.nc.L2::128B", t, mask, out, n);
    run<4>("ld.global.nc.L1::no_allocate", t, mask, out, n);
    run<5>("ld.global.cg", t, mask, out, n);
    run<6>("ld.global.cs", t, mask, out, n);
    run<7>("ld.global.nc.L1::evict_first", t, mask, out, n);
    run<8>("ld.global.lu", t, mask, out, n);
  }
  return 0;
}

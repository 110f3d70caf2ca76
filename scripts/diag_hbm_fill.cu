// Write-only HBM bandwidth on B200: what is the ceiling a kernel that stores every byte of a
// 1 GiB buffer exactly once (the table generator) can reach?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/build/diag_hbm_fill scripts/diag_hbm_fill.cu
//   gpurun -- ./scripts/build/diag_hbm_fill > gpurun_out/hbm_fill.txt
//
// Variants: cudaMemsetAsync; grid-stride fill kernels with 128-bit (st.global.v4.b32) and
// 256-bit (st.global.v8.b32, sm_100) stores, default / .cs / .wt cache hints, several grid
// shapes; shared memory -> global bulk copies (cp.async.bulk.global.shared::cta, the TMA
// store path) with one or more bulk copies in flight per CTA; and, for scale, a copy
// (read + write) and a read-only sum.  Prints GB/s = bytes written (or moved) / time.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e = (x);                                                        \
    if (e != cudaSuccess) {                                                     \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                   \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

constexpr size_t kBytes = 1ull << 30;

enum Hint { kDefault, kCs, kWt, kNoAlloc };

template <int HINT>
__device__ __forceinline__ void st128(uint4* p, uint4 v) {
  if (HINT == kCs)
    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
  else if (HINT == kWt)
    asm volatile("st.global.wt.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
  else if (HINT == kNoAlloc)
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
  else
    *p = v;
}

// thread-contiguous 16 B, warp-contiguous 512 B, grid-stride
template <int HINT>
__global__ void fill128(uint4* out, size_t n16, uint32_t seed) {
  const uint4 v = make_uint4(seed, seed + 1, seed + 2, seed + 3);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
       i += (size_t)gridDim.x * blockDim.x)
    st128<HINT>(out + i, v);
}

// 256-bit stores: each thread 32 B, warp 1 KiB
__global__ void fill256(uint4* out, size_t n32, uint32_t seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32;
       i += (size_t)gridDim.x * blockDim.x) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out + 2 * i),
                 "r"(seed), "r"(seed + 1), "r"(seed + 2), "r"(seed + 3), "r"(seed + 4),
                 "r"(seed + 5), "r"(seed + 6), "r"(seed + 7)
                 : "memory");
  }
}

// Like the generator: a block owns a contiguous chunk, each warp writes 512-byte rows that
// are `row_stride` bytes apart (one STG.128 per lane and row), 8 rows per pass.
__global__ void fill_rows(uint4* out, size_t row_stride16, int rows_per_block, uint32_t seed) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // warp w of block b writes columns [32 * (b % cols) ...]: emulate 16 KiB rows (HGR: 2^14
  // entries x 2 B = 32 KiB per row; here row_stride16 * 16 bytes)
  const size_t col_blocks = row_stride16 / 32;   // 512-byte column blocks per row
  const size_t cb = (size_t)blockIdx.x * nw + warp;
  const size_t col = cb % col_blocks, rblk = cb / col_blocks;
  uint4* p = out + (rblk * rows_per_block) * row_stride16 + col * 32 + lane;
  const uint4 v = make_uint4(seed, seed + 1, seed + 2, seed + 3);
  for (int r = 0; r < rows_per_block; ++r) p[(size_t)r * row_stride16] = v;
}

// shared -> global bulk copies (TMA store path, no tensor map needed)
template <int STAGES>
__global__ void fill_bulk(uint8_t* out, size_t chunk, size_t n_chunks, uint32_t seed) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (size_t i = threadIdx.x; i < (chunk * STAGES) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = seed + (uint32_t)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = 0;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++k) {
      const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem + (k % STAGES) * chunk);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                       out + c * chunk),
                   "r"(s), "r"((uint32_t)chunk)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // keep at most STAGES groups in flight (the source buffers are never rewritten here,
      // but a real producer would have to wait before refilling a stage)
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void copy128(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
       i += (size_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}

__global__ void sum128(const uint4* __restrict__ in, size_t n16, unsigned long long* res) {
  uint32_t acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
       i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = in[i];
    acc += v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) atomicAdd(res, 1ull);
}

template <typename F>
static float time_ms(F&& launch, int reps = 40) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 5; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

static void report(const char* name, float ms, double bytes) {
  printf("%-64s %8.4f ms  %8.1f GB/s\n", name, ms, bytes / (ms * 1e-3) / 1e9);
  fflush(stdout);
}

int main() {
  uint8_t *x, *y;
  unsigned long long* res;
  CK(cudaMalloc(&x, kBytes));
  CK(cudaMalloc(&y, kBytes));
  CK(cudaMalloc(&res, 8));
  CK(cudaMemset(y, 1, kBytes));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("# %s, %d SMs; 1 GiB buffer; GB/s = 1e9 bytes/s\n", prop.name, sms);
  const size_t n16 = kBytes / 16;
  char name[160];

  report("cudaMemsetAsync", time_ms([&] { CK(cudaMemsetAsync(x, 7, kBytes)); }), (double)kBytes);
  for (int bps : {2, 4, 8, 16}) {
    for (int threads : {256, 512}) {
      const int grid = sms * bps;
      snprintf(name, sizeof name, "fill128 default, grid %d x %d threads", grid, threads);
      report(name, time_ms([&] { fill128<kDefault><<<grid, threads>>>((uint4*)x, n16, 3); }),
             (double)kBytes);
    }
  }
  {
    const int grid = (int)(n16 / 256);
    report("fill128 default, one 16 B store per thread (grid = n/256)",
           time_ms([&] { fill128<kDefault><<<grid, 256>>>((uint4*)x, n16, 3); }), (double)kBytes);
  }
  report("fill128 .cs, grid 8/SM x 256",
         time_ms([&] { fill128<kCs><<<sms * 8, 256>>>((uint4*)x, n16, 3); }), (double)kBytes);
  report("fill128 .wt, grid 8/SM x 256",
         time_ms([&] { fill128<kWt><<<sms * 8, 256>>>((uint4*)x, n16, 3); }), (double)kBytes);
  report("fill128 L1::no_allocate, grid 8/SM x 256",
         time_ms([&] { fill128<kNoAlloc><<<sms * 8, 256>>>((uint4*)x, n16, 3); }),
         (double)kBytes);
  for (int bps : {4, 8, 16})
    {
      snprintf(name, sizeof name, "fill256 (st.global.v8.b32), grid %d/SM x 256", bps);
      report(name, time_ms([&] { fill256<<<sms * bps, 256>>>((uint4*)x, n16 / 2, 3); }),
             (double)kBytes);
    }
  // generator-like: 32 KiB rows (HGR), a warp writes 512 B of each of 8 / 128 rows
  for (int rows : {8, 128}) {
    const size_t row16 = 32768 / 16;
    const size_t n_rows = kBytes / 32768;
    const size_t warps = (n_rows / rows) * (row16 / 32);
    snprintf(name, sizeof name, "fill_rows: warp = 512 B x %d rows of 32 KiB, 128 thr/block", rows);
    report(name, time_ms([&] {
             fill_rows<<<(unsigned)(warps / 4), 128>>>((uint4*)x, row16, rows, 3);
           }),
           (double)kBytes);
  }
  // bulk (TMA) stores from shared memory
  for (size_t chunk : {(size_t)4096, (size_t)16384, (size_t)32768}) {
    const size_t n_chunks = kBytes / chunk;
    for (int bps : {1, 2, 4}) {
      if (chunk * 2 * bps > 200 * 1024) continue;
      CK(cudaFuncSetAttribute(fill_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(chunk * 2)));
      snprintf(name, sizeof name, "bulk smem->global, %zu B chunks, 2 in flight, %d CTA/SM", chunk, bps);
      report(name, time_ms([&] {
               fill_bulk<2><<<sms * bps, 128, chunk * 2>>>(x, chunk, n_chunks, 3);
             }),
             (double)kBytes);
    }
    if (chunk * 4 <= 200 * 1024) {
      CK(cudaFuncSetAttribute(fill_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(chunk * 4)));
      snprintf(name, sizeof name, "bulk smem->global, %zu B chunks, 4 in flight, 1 CTA/SM", chunk);
      report(name, time_ms([&] {
               fill_bulk<4><<<sms, 128, chunk * 4>>>(x, chunk, n_chunks, 3);
             }),
             (double)kBytes);
    }
  }
  report("copy128 (read 1 GiB + write 1 GiB), grid 8/SM x 256",
         time_ms([&] { copy128<<<sms * 8, 256>>>((const uint4*)y, (uint4*)x, n16); }),
         2.0 * kBytes);
  report("cudaMemcpyAsync D2D (read + write)",
         time_ms([&] { CK(cudaMemcpyAsync(x, y, kBytes, cudaMemcpyDeviceToDevice)); }),
         2.0 * kBytes);
  report("sum128 (read 1 GiB), grid 8/SM x 256",
         time_ms([&] { sum128<<<sms * 8, 256>>>((const uint4*)y, n16, res); }), (double)kBytes);
  return 0;
}

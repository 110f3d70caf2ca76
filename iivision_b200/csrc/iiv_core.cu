// Error reporting, version and class-constant queries of the C ABI.
#include <cstdarg>
#include <cstdio>

#include "iiv_common.cuh"

namespace iiv {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return (int)e;
}

}  // namespace iiv

extern "C" const char* iiv_last_error(void) { return iiv::g_error; }

extern "C" int iiv_version(void) { return 100; }

extern "C" int iiv_mode_info(int mode, int* masked_bits, int* masked_dots,
                             int* n_offsets, int* phases4) {
  using namespace iiv;
  IIV_REQUIRE(mode == IIV_MODE_HGR || mode == IIV_MODE_DHGR, "bad mode %d", mode);
  if (mode == IIV_MODE_HGR) {
    using M = Mode<IIV_MODE_HGR>;
    if (masked_bits) *masked_bits = M::kBits;
    if (masked_dots) *masked_dots = M::kDots;
    if (n_offsets) *n_offsets = M::kOffsets;
    if (phases4)
      for (int o = 0; o < M::kOffsets; ++o) phases4[o] = M::phase(o);
  } else {
    using M = Mode<IIV_MODE_DHGR>;
    if (masked_bits) *masked_bits = M::kBits;
    if (masked_dots) *masked_dots = M::kDots;
    if (n_offsets) *n_offsets = M::kOffsets;
    if (phases4)
      for (int o = 0; o < M::kOffsets; ++o) phases4[o] = M::phase(o);
  }
  return 0;
}

// The scorer's table reads are 2-byte gathers spread uniformly over a 512 MiB / 1 GiB table:
// every miss should cost one 32-byte sector, not the default wider fetch.
extern "C" int iiv_set_l2_fetch_granularity(size_t bytes) {
  using namespace iiv;
  IIV_REQUIRE(bytes == 32 || bytes == 64 || bytes == 128, "granularity must be 32, 64 or 128");
  IIV_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, bytes));
  return 0;
}

extern "C" size_t iiv_get_l2_fetch_granularity(void) {
  size_t v = 0;
  if (cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity) != cudaSuccess) return 0;
  return v;
}

// Write-only bandwidth probes for bench.py's roofline object: what a kernel that stores every
// byte of a buffer exactly once (the table generator) can reach at best on this device.
namespace iiv {
namespace {
__global__ void __launch_bounds__(256) fill_probe_kernel(uint4* out, size_t n16) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;   // one 16-byte store per thread
  if (i < n16) out[i] = make_uint4(0u, 1u, 2u, 3u);
}
}  // namespace
}  // namespace iiv

extern "C" int iiv_fill_probe(void* d_buf, size_t bytes, int variant, void* stream) {
  using namespace iiv;
  IIV_REQUIRE(d_buf && bytes % 16 == 0 && (variant == 0 || variant == 1), "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (variant == 0) {
    IIV_CUDA(cudaMemsetAsync(d_buf, 0, bytes, st));
  } else {
    const size_t n16 = bytes / 16;
    fill_probe_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>((uint4*)d_buf, n16);
    IIV_LAUNCH_CHECK("fill_probe_kernel");
  }
  return 0;
}

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

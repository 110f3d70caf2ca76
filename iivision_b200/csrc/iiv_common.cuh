// Shared declarations for libiivision_b200 (sm_100a only).
//
// Bit layouts follow the reference's packed-bitmap classes:
//   HGR  (transcoder/screen.py:565-645): 22-bit words  ffFbbbbbbbBAaaaaaaaHhh
//   DHGR (transcoder/screen.py:822-919): 34-bit words  fff G..A hhh
// one word per pair of page offsets (2c, 2c+1), uint64[32][128] per screen.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/iivision_b200.h"

namespace iiv {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define IIV_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t e__ = (expr);                                 \
    if (e__ != cudaSuccess) return iiv::cuda_fail(e__, #expr); \
  } while (0)

#define IIV_LAUNCH_CHECK(name)                                   \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return iiv::cuda_fail(e__, name);    \
  } while (0)

#define IIV_REQUIRE(cond, ...)       \
  do {                               \
    if (!(cond)) {                   \
      iiv::set_error(__VA_ARGS__);   \
      return IIV_E_BADARG;           \
    }                                \
  } while (0)

template <int MODE>
struct Mode;

template <>
struct Mode<IIV_MODE_HGR> {
  static constexpr int kBits = 14;      // MASKED_BITS  screen.py:617
  static constexpr int kDots = 18;      // MASKED_DOTS  screen.py:626
  static constexpr int kOffsets = 2;    // len(BYTE_MASKS) screen.py:632-636
  static constexpr int kHeader = 3, kBody = 16, kFooter = 3;  // :609-612
  static constexpr int kContents = 256;
  __host__ __device__ static constexpr int phase(int o) {      // :645
    return o == 0 ? 1 : 3;
  }
  __host__ __device__ static constexpr int shift(int o) { return 8 * o; }
};

template <>
struct Mode<IIV_MODE_DHGR> {
  static constexpr int kBits = 13;      // screen.py:887
  static constexpr int kDots = 10;      // screen.py:891
  static constexpr int kOffsets = 4;    // screen.py:894-908
  static constexpr int kHeader = 3, kBody = 28, kFooter = 3;  // :882-884
  static constexpr int kContents = 128;
  __host__ __device__ static constexpr int phase(int o) {      // :919
    return o == 0 ? 1 : o == 1 ? 0 : o == 2 ? 3 : 2;
  }
  __host__ __device__ static constexpr int shift(int o) { return 7 * o; }
};

// Bitmap.mask_and_shift_data (screen.py:369-378).
template <int MODE>
__host__ __device__ __forceinline__ uint32_t mask_shift(uint64_t w, int o) {
  return (uint32_t)(w >> Mode<MODE>::shift(o)) &
         ((1u << Mode<MODE>::kBits) - 1u);
}

// masked_update (screen.py:791-816 HGR, :992-1007 DHGR).
template <int MODE>
__host__ __device__ __forceinline__ uint64_t masked_update(int o, uint64_t old,
                                                           uint32_t v) {
  if (MODE == IIV_MODE_HGR) {
    if (o == 0) return (old & ~(uint64_t)(0xffu << 3)) ^ ((uint64_t)(v & 0xffu) << 3);
    const uint32_t rot = ((v & 0x7fu) << 1) ^ ((v & 0x80u) >> 7);
    return (old & ~(uint64_t)(0xffu << 11)) ^ ((uint64_t)rot << 11);
  } else {
    const int sh = 7 * o + 3;
    return (old & ~((uint64_t)0x7f << sh)) ^ ((uint64_t)(v & 0x7fu) << sh);
  }
}

// _make_header (screen.py:650-661 HGR, :921-924 DHGR): header bits the column
// contributes to its right-hand neighbour.
template <int MODE>
__host__ __device__ __forceinline__ uint64_t header_of(uint64_t col) {
  if (MODE == IIV_MODE_HGR)
    return ((col & (1ull << 11)) >> 9) ^ ((col & (3ull << 17)) >> 17);
  return (col >> 28) & 7ull;
}

// _make_footer (screen.py:679-690 HGR, :949-952 DHGR): footer bits the column
// contributes to its left-hand neighbour.
template <int MODE>
__host__ __device__ __forceinline__ uint64_t footer_of(uint64_t col) {
  if (MODE == IIV_MODE_HGR)
    return (((col & (1ull << 10)) >> 10) ^ ((col & (3ull << 3)) >> 2)) << 19;
  return (col & (7ull << 3)) << 28;
}

template <int MODE>
__host__ __device__ __forceinline__ uint64_t keep_low_mask() {
  return (1ull << (Mode<MODE>::kHeader + Mode<MODE>::kBody)) - 1ull;
}
template <int MODE>
__host__ __device__ __forceinline__ uint64_t keep_high_mask() {
  return ((1ull << (Mode<MODE>::kBody + Mode<MODE>::kFooter)) - 1ull)
         << Mode<MODE>::kHeader;
}

// byte_offset (screen.py:692-700 HGR, :954-968 DHGR).
template <int MODE>
__host__ __device__ __forceinline__ int byte_offset(int page_offset, int is_aux) {
  const int odd = page_offset & 1;
  if (MODE == IIV_MODE_HGR) return odd;
  return is_aux ? (odd ? 2 : 0) : (odd ? 3 : 1);
}

// SCREEN_HOLES (screen.py:41-62): offsets 120..127 and 248..255 of every page.
__host__ __device__ __forceinline__ bool is_hole(int page_offset) {
  return (page_offset & 127) >= 120;
}

// Body of one packed column (screen.py:663-677 HGR, :926-947 DHGR).
template <int MODE>
__host__ __device__ __forceinline__ uint64_t body_of(uint32_t main_even,
                                                     uint32_t main_odd,
                                                     uint32_t aux_even,
                                                     uint32_t aux_odd) {
  if (MODE == IIV_MODE_HGR)
    return ((uint64_t)main_even << 3) + ((uint64_t)(main_odd & 0x7fu) << 12) +
           ((uint64_t)(main_odd & 0x80u) << 4);
  return ((uint64_t)(aux_even & 0x7fu) << 3) +
         ((uint64_t)(main_even & 0x7fu) << 10) +
         ((uint64_t)(aux_odd & 0x7fu) << 17) +
         ((uint64_t)(main_odd & 0x7fu) << 24);
}

// HGRBitmap._double_pixels (screen.py:710-739).
__host__ __device__ __forceinline__ uint32_t double_pixels(uint32_t x7) {
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) out |= ((x7 >> k) & 1u) * (3u << (2 * k));
  out |= ((x7 >> 6) & 1u) * (7u << 12);
  return out;
}

// Bitmap.to_dots (screen.py:741-789 HGR; :982-990 DHGR is the identity).
template <int MODE>
__host__ __device__ __forceinline__ uint32_t to_dots(uint32_t v, int o) {
  if (MODE == IIV_MODE_DHGR) return v;
  const uint32_t h = (v & 7u) << 5;
  const uint32_t hp = (h & 0x80u) >> 7;
  uint32_t res = double_pixels(h & 0x7fu) >> (11 - hp);
  uint32_t b, bp;
  if (o == 0) {
    b = (v >> 3) & 0xffu;
    bp = (b & 0x80u) >> 7;
  } else {
    bp = (v >> 3) & 1u;
    b = ((v >> 4) & 0x7fu) ^ (bp << 7);
  }
  res &= ~(0x3fffu << (3 + bp));
  res ^= double_pixels(b & 0x7fu) << (3 + bp);
  const uint32_t f = ((v >> 12) & 3u) ^ (((v >> 11) & 1u) << 7);
  const uint32_t fp = (f & 0x80u) >> 7;
  res &= ~(0xfu << (17 + fp));
  res ^= double_pixels(f & 0x7fu) << (17 + fp);
  return res & ((1u << 21) - 1u);
}

// colours.py:100-134: pixel t = rol4((dots >> t) & 15, (phase + t) & 3).
__host__ __device__ __forceinline__ uint32_t nominal_pixel(uint32_t dots, int t,
                                                           int init_phase) {
  const uint32_t w = (dots >> t) & 0xfu;
  const int ph = (init_phase + t) & 3;
  return ((w << ph) | (w >> (4 - ph))) & 0xfu;
}

// One entry of an edit-distance table (the scorer's random 2-byte gathers, screen.py:441,
// :486).  Read-only path; the L2::64B hint measured 5 % more gathers/s than a plain
// ld.global.nc on uniformly random indices over the 512 MiB DHGR table
// (profiles/r02_gather_flavours.txt) -- DRAM moves ~107 B per missing gather either way.
__device__ __forceinline__ uint32_t ldg_table(const uint16_t* p) {
  uint16_t v;
  asm("ld.global.nc.L2::64B.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}

// ---- bit windows of a masked value (table generator's halves, scorer's segments) ----
// Compile-time bit window: ext gathers the window's bits of v into a dense index (lowest bit
// first), dep scatters an index back; one shift + mask per run of the window.
__host__ __device__ constexpr int ctz_c(uint32_t x) {
  int n = 0;
  while (n < 32 && !((x >> n) & 1u)) ++n;
  return n;
}
template <uint32_t MASK, int OUT = 0>
struct Ext {
  static constexpr int lo = ctz_c(MASK);
  static constexpr int len = ctz_c(~(MASK >> lo));
  static constexpr uint32_t rest = MASK & ~(((1u << len) - 1u) << lo);
  static constexpr int count = len + Ext<rest, OUT + len>::count;
  __host__ __device__ static __forceinline__ constexpr uint32_t ext(uint32_t v) {
    return (((v >> lo) & ((1u << len) - 1u)) << OUT) | Ext<rest, OUT + len>::ext(v);
  }
  __host__ __device__ static __forceinline__ constexpr uint32_t dep(uint32_t x) {
    return (((x >> OUT) & ((1u << len) - 1u)) << lo) | Ext<rest, OUT + len>::dep(x);
  }
};
template <int OUT>
struct Ext<0u, OUT> {
  static constexpr int count = 0;
  __host__ __device__ static __forceinline__ constexpr uint32_t ext(uint32_t) { return 0u; }
  __host__ __device__ static __forceinline__ constexpr uint32_t dep(uint32_t) { return 0u; }
};

// ---- MT19937 (numpy legacy RandomState and CPython random share it) --------
__host__ __device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

__host__ __device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t nxt,
                                                    uint32_t far) {
  const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

}  // namespace iiv

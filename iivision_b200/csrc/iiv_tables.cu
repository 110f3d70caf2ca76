// Path 1: edit-distance table generation
// (reference transcoder/make_data_tables.py:111-174 compute_edit_distance with
//  edit_distance :92-108 and colours.py:100-148 inlined).
//
// An entry T[o][(i << bits) + j] is the weighted Damerau-Levenshtein distance
// between the n-pixel nominal-colour strings of masked values i and j at byte
// offset o.  Insert/delete cost 1e5 (make_data_tables.py:35-36) while any
// substitution-only alignment costs <= n * 255, so the optimal alignment never
// inserts or deletes and the (n+2)^2 DP of weighted_levenshtein.dam_lev collapses
// to the chain
//     D_t = min(D_{t-1} + S[a_t][b_t],  D_{t-2} + 1 if a_{t-1}==b_t && a_t==b_{t-1})
// (tests/test_oracle_tables.py checks the collapse against the
// restated full DP).  All arithmetic is exact small-integer work.
#include "iiv_common.cuh"

namespace iiv {
namespace {

// Pixel strings, 4 bits per pixel, pixel t at bits [4t, 4t+4).  HGR needs 72
// bits: .x/.y = low 64, .z = pixels 16..17.  Filled by pixel_prologue on the
// stream of each generate call; the content depends on the mode only.
__device__ uint4 g_pix_hgr[2][1 << 14];
__device__ uint2 g_pix_dhgr[4][1 << 13];

struct Lut {
  uint8_t s[256];
};
struct Lut32 {
  int32_t s[256];
};

// Destination tables of one generate call: the local table, or every rank's
// peer-mapped table (fused generate + all-gather over NVLink), or one NVSwitch
// multicast mapping.
constexpr int kMaxDests = 8;
struct Dests {
  uint16_t* p[kMaxDests];
  int n;
};

template <int MODE>
__global__ void pixel_prologue() {
  using M = Mode<MODE>;
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const int o = blockIdx.y;
  if (v >= (1u << M::kBits)) return;
  const uint32_t dots = to_dots<MODE>(v, o);
  uint64_t lo = 0;
  uint32_t hi = 0;
#pragma unroll
  for (int t = 0; t < M::kDots; ++t) {
    const uint64_t p = nominal_pixel(dots, t, M::phase(o));
    if (t < 16)
      lo |= p << (4 * t);
    else
      hi |= (uint32_t)p << (4 * (t - 16));
  }
  if (MODE == IIV_MODE_HGR)
    g_pix_hgr[o][v] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), hi, 0u);
  else
    g_pix_dhgr[o][v] = make_uint2((uint32_t)lo, (uint32_t)(lo >> 32));
}

template <int MODE>
__global__ void dots_kernel(uint32_t* out) {
  using M = Mode<MODE>;
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const int o = blockIdx.y;
  if (v < (1u << M::kBits)) out[((size_t)o << M::kBits) + v] = to_dots<MODE>(v, o);
}

template <int MODE>
__global__ void pixel_strings_kernel(uint8_t* out) {
  using M = Mode<MODE>;
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const int o = blockIdx.y;
  if (v >= (1u << M::kBits)) return;
  const uint32_t dots = to_dots<MODE>(v, o);
  uint8_t* p = out + (((size_t)o << M::kBits) + v) * M::kDots;
  for (int t = 0; t < M::kDots; ++t)
    p[t] = (uint8_t)nominal_pixel(dots, t, M::phase(o));
}

template <int MODE>
__device__ __forceinline__ void load_pixels(int o, uint32_t v, uint64_t& lo,
                                            uint32_t& hi) {
  if (MODE == IIV_MODE_HGR) {
    const uint4 q = g_pix_hgr[o][v];
    lo = ((uint64_t)q.y << 32) | q.x;
    hi = q.z;
  } else {
    const uint2 q = g_pix_dhgr[o][v];
    lo = ((uint64_t)q.y << 32) | q.x;
    hi = 0;
  }
}

__device__ __forceinline__ uint32_t pixel_at(uint64_t lo, uint32_t hi, int t) {
  return t < 16 ? (uint32_t)(lo >> (4 * t)) & 15u : (hi >> (4 * (t - 16))) & 15u;
}

// ---- ALGO_CHAIN: one independent recurrence per entry ------------------------
// Block = 256 threads covering 2048 consecutive j of one row i and offset o;
// each thread produces 8 consecutive entries -> one 128-bit store.
constexpr int kChainPerThread = 8;
constexpr int kChainThreads = 256;

template <int MODE>
__global__ void __launch_bounds__(kChainThreads)
chain_kernel(Lut lut, Dests dests, uint32_t row_begin, int triangular) {
  using M = Mode<MODE>;
  __shared__ uint8_t S[256];
  S[threadIdx.x] = lut.s[threadIdx.x];
  __syncthreads();

  const int o = blockIdx.z;
  const uint32_t i = row_begin + blockIdx.y;
  const uint32_t j0 =
      (blockIdx.x * kChainThreads + threadIdx.x) * kChainPerThread;
  const size_t at = ((size_t)o << (2 * M::kBits)) + ((size_t)i << M::kBits) + j0;
  if (triangular && j0 >= i) {
    for (int d = 0; d < dests.n; ++d)
      *reinterpret_cast<uint4*>(dests.p[d] + at) = make_uint4(0, 0, 0, 0);
    return;
  }
  uint64_t alo;
  uint32_t ahi;
  load_pixels<MODE>(o, i, alo, ahi);

  uint32_t res[kChainPerThread];
#pragma unroll
  for (int k = 0; k < kChainPerThread; ++k) {
    uint64_t blo;
    uint32_t bhi;
    load_pixels<MODE>(o, j0 + k, blo, bhi);
    uint32_t d2 = 0, d1 = 0, pa = 0, pb = 0;
#pragma unroll
    for (int t = 0; t < M::kDots; ++t) {
      const uint32_t a = pixel_at(alo, ahi, t);
      const uint32_t b = pixel_at(blo, bhi, t);
      uint32_t cur = d1 + S[a * 16 + b];
      if (t >= 1 && pa == b && a == pb) cur = min(cur, d2 + 1u);
      d2 = d1;
      d1 = cur;
      pa = a;
      pb = b;
    }
    res[k] = (triangular && j0 + k >= i) ? 0u : d1;
  }
  uint4 v;
  v.x = res[0] | (res[1] << 16);
  v.y = res[2] | (res[3] << 16);
  v.z = res[4] | (res[5] << 16);
  v.w = res[6] | (res[7] << 16);
  for (int d = 0; d < dests.n; ++d) *reinterpret_cast<uint4*>(dests.p[d] + at) = v;
}

// edit_distance (make_data_tables.py:92-108) for explicit pixel strings: pairs of
// `len` nibble-valued pixels, one thread per pair.
__global__ void string_distance_kernel(Lut32 lut, const uint8_t* __restrict__ a,
                                       const uint8_t* __restrict__ b, int n_pairs,
                                       int len, int32_t* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_pairs) return;
  const uint8_t* pa = a + (size_t)k * len;
  const uint8_t* pb = b + (size_t)k * len;
  int32_t d2 = 0, d1 = 0;
  for (int t = 0; t < len; ++t) {
    int32_t cur = d1 + lut.s[(pa[t] & 15) * 16 + (pb[t] & 15)];
    if (t >= 1 && pa[t - 1] == pb[t] && pa[t] == pb[t - 1]) cur = min(cur, d2 + 1);
    d2 = d1;
    d1 = cur;
  }
  out[k] = d1;
}

// ---- Bitmap.edit_distances transform (screen.py:358-365) ----------------------
// new[y] = old[y] + old[T(y)], T swapping the two index halves.  One block owns
// a 32x32 tile pair {(bi,bj),(bj,bi)}, bi >= bj, so the update is race-free.
template <int MODE>
__global__ void __launch_bounds__(256)
symmetrise_kernel(uint16_t* __restrict__ table) {
  using M = Mode<MODE>;
  constexpr uint32_t NT = (1u << M::kBits) / 32;
  __shared__ uint16_t A[32][33], B[32][33];
  // Linear block index -> (bi, bj) with bj <= bi.
  const uint32_t lin = blockIdx.x;
  uint32_t bi = (uint32_t)((sqrtf(8.0f * (float)lin + 1.0f) - 1.0f) * 0.5f);
  while ((bi + 1) * (bi + 2) / 2 <= lin) ++bi;
  while (bi * (bi + 1) / 2 > lin) --bi;
  const uint32_t bj = lin - bi * (bi + 1) / 2;
  if (bi >= NT) return;
  uint16_t* t = table + ((size_t)blockIdx.y << (2 * M::kBits));
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    A[r][tx] = t[((size_t)(bi * 32 + r) << M::kBits) + bj * 32 + tx];
    B[r][tx] = t[((size_t)(bj * 32 + r) << M::kBits) + bi * 32 + tx];
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    t[((size_t)(bi * 32 + r) << M::kBits) + bj * 32 + tx] =
        (uint16_t)(A[r][tx] + B[tx][r]);
    if (bi != bj)
      t[((size_t)(bj * 32 + r) << M::kBits) + bi * 32 + tx] =
          (uint16_t)(B[r][tx] + A[tx][r]);
  }
}

template <int MODE>
int generate(const Lut& lut, const Dests& dests, uint32_t row_begin,
             uint32_t row_end, int layout, int algo, cudaStream_t st) {
  using M = Mode<MODE>;
  constexpr uint32_t N = 1u << M::kBits;
  IIV_REQUIRE(row_begin <= row_end && row_end <= N, "bad row range [%u,%u)",
              row_begin, row_end);
  if (row_begin == row_end) return 0;
  pixel_prologue<MODE><<<dim3(N / 256, M::kOffsets), 256, 0, st>>>();
  IIV_LAUNCH_CHECK("pixel_prologue");
  (void)algo;
  const uint32_t rows = row_end - row_begin;
  // gridDim.y <= 65535 holds: rows <= 16384.
  dim3 grid(N / (kChainThreads * kChainPerThread), rows, M::kOffsets);
  chain_kernel<MODE><<<grid, kChainThreads, 0, st>>>(
      lut, dests, row_begin, layout == IIV_LAYOUT_TRIANGULAR);
  IIV_LAUNCH_CHECK("chain_kernel");
  return 0;
}

}  // namespace
}  // namespace iiv

using namespace iiv;

extern "C" int iiv_all_dots(int mode, uint32_t* d_dots, void* stream) {
  IIV_REQUIRE(d_dots, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == IIV_MODE_HGR)
    dots_kernel<IIV_MODE_HGR><<<dim3(64, 2), 256, 0, st>>>(d_dots);
  else if (mode == IIV_MODE_DHGR)
    dots_kernel<IIV_MODE_DHGR><<<dim3(32, 4), 256, 0, st>>>(d_dots);
  else
    IIV_REQUIRE(false, "bad mode %d", mode);
  IIV_LAUNCH_CHECK("dots_kernel");
  return 0;
}

extern "C" int iiv_all_pixel_strings(int mode, uint8_t* d_pix, void* stream) {
  IIV_REQUIRE(d_pix, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == IIV_MODE_HGR)
    pixel_strings_kernel<IIV_MODE_HGR><<<dim3(64, 2), 256, 0, st>>>(d_pix);
  else if (mode == IIV_MODE_DHGR)
    pixel_strings_kernel<IIV_MODE_DHGR><<<dim3(32, 4), 256, 0, st>>>(d_pix);
  else
    IIV_REQUIRE(false, "bad mode %d", mode);
  IIV_LAUNCH_CHECK("pixel_strings_kernel");
  return 0;
}

static int make_lut(const int32_t* h_lut, Lut* lut) {
  for (int k = 0; k < 256; ++k) {
    if (h_lut[k] < 0 || h_lut[k] > 255) {
      set_error("substitution cost %d at [%d][%d] outside 0..255", h_lut[k],
                k >> 4, k & 15);
      return IIV_E_OVERFLOW;
    }
    lut->s[k] = (uint8_t)h_lut[k];
  }
  return 0;
}

static int generate_any(int mode, const int32_t* h_lut, const Dests& dests,
                        uint32_t row_begin, uint32_t row_end, int layout, int algo,
                        void* stream) {
  IIV_REQUIRE(layout == IIV_LAYOUT_TRIANGULAR || layout == IIV_LAYOUT_SYMMETRIC,
              "bad layout %d", layout);
  Lut lut;
  const int rc = make_lut(h_lut, &lut);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == IIV_MODE_HGR)
    return generate<IIV_MODE_HGR>(lut, dests, row_begin, row_end, layout, algo, st);
  if (mode == IIV_MODE_DHGR)
    return generate<IIV_MODE_DHGR>(lut, dests, row_begin, row_end, layout, algo, st);
  IIV_REQUIRE(false, "bad mode %d", mode);
}

extern "C" int iiv_table_generate(int mode, const int32_t* h_lut,
                                  uint16_t* d_table, uint32_t row_begin,
                                  uint32_t row_end, int layout, int algo,
                                  void* stream) {
  IIV_REQUIRE(h_lut && d_table, "null pointer");
  Dests dests;
  dests.n = 1;
  dests.p[0] = d_table;
  return generate_any(mode, h_lut, dests, row_begin, row_end, layout, algo, stream);
}

extern "C" int iiv_table_generate_scatter(int mode, const int32_t* h_lut,
                                          uint16_t* const* h_peer_tables,
                                          int n_ranks, int rank,
                                          uint16_t* d_multicast_table,
                                          uint32_t row_begin, uint32_t row_end,
                                          int layout, void* stream) {
  IIV_REQUIRE(h_lut, "null pointer");
  IIV_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxDests && rank >= 0 && rank < n_ranks,
              "bad rank %d of %d", rank, n_ranks);
  Dests dests;
  if (d_multicast_table) {
    dests.n = 1;
    dests.p[0] = d_multicast_table;
  } else {
    IIV_REQUIRE(h_peer_tables, "null pointer");
    dests.n = n_ranks;
    // own table first, then peers in ring order so ranks do not all hit the same
    // peer at the same time
    for (int k = 0; k < n_ranks; ++k) {
      dests.p[k] = h_peer_tables[(rank + k) % n_ranks];
      IIV_REQUIRE(dests.p[k], "null peer table %d", (rank + k) % n_ranks);
    }
  }
  return generate_any(mode, h_lut, dests, row_begin, row_end, layout, IIV_ALGO_AUTO,
                      stream);
}

extern "C" int iiv_string_distance(const int32_t* h_lut, const uint8_t* d_a,
                                   const uint8_t* d_b, int n_pairs, int len,
                                   int32_t* d_out, void* stream) {
  IIV_REQUIRE(h_lut && d_a && d_b && d_out && n_pairs >= 0 && len >= 0, "bad argument");
  // edit_distance(error=True) uses 5x costs (make_data_tables.py:85-86), so the
  // explicit-string entry point takes any cost that keeps the sum in int32.
  Lut32 lut;
  for (int k = 0; k < 256; ++k) {
    if (h_lut[k] < 0 || (int64_t)h_lut[k] * len > 0x7fffffff) {
      set_error("substitution cost %d at [%d][%d] out of range", h_lut[k], k >> 4, k & 15);
      return IIV_E_OVERFLOW;
    }
    lut.s[k] = h_lut[k];
  }
  if (n_pairs == 0) return 0;
  string_distance_kernel<<<(n_pairs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      lut, d_a, d_b, n_pairs, len, d_out);
  IIV_LAUNCH_CHECK("string_distance_kernel");
  return 0;
}

extern "C" int iiv_table_symmetrise(int mode, uint16_t* d_table, void* stream) {
  IIV_REQUIRE(d_table, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == IIV_MODE_HGR) {
    constexpr uint32_t NT = (1u << 14) / 32;
    symmetrise_kernel<IIV_MODE_HGR>
        <<<dim3(NT * (NT + 1) / 2, 2), 256, 0, st>>>(d_table);
  } else if (mode == IIV_MODE_DHGR) {
    constexpr uint32_t NT = (1u << 13) / 32;
    symmetrise_kernel<IIV_MODE_DHGR>
        <<<dim3(NT * (NT + 1) / 2, 4), 256, 0, st>>>(d_table);
  } else {
    IIV_REQUIRE(false, "bad mode %d", mode);
  }
  IIV_LAUNCH_CHECK("symmetrise_kernel");
  return 0;
}

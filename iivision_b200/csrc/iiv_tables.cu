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
  int multicast;  // p[0] is an NVSwitch multicast mapping: one multimem.st reaches all
};

// 16-byte store through a multicast (multimem) address: the switch replicates it
// into every member GPU's copy of the buffer.
__device__ __forceinline__ void multimem_st_v4(uint16_t* addr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr),
               "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
               "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}

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

// ---- ALGO_TREE: shared-suffix tree over 8x8 tiles ---------------------------------
// The chain can be run from the last pixel backwards,
//     F_t = min(F_{t+1} + S[a_t][b_t],  F_{t+2} + 1 if a_t==b_{t+1} && a_{t+1}==b_t),
// F_n = 0, entry = F_0 (same set of tilings by singles and swapped pairs as the
// forward form).  The low three bits of a masked value only reach the first
// kLeaf pixels (HGR: bit 1 -> pixels 0..3, bits 0 and 2 -> pixels 0..1; DHGR: bit k
// -> pixels 0..k; tests/test_oracle_tables.py::test_low_bits_reach_leaf_pixels_only),
// so an 8x8 tile of entries (8 consecutive i) x (8 consecutive j) shares
// F_kLeaf, F_kLeaf+1 and only the leaf pixels are walked per entry, as a tree over
// the bits each pixel depends on:
//   HGR : 14 shared steps + 4 x (pixels 3,2) + 64 x (pixels 1,0) = 150 steps / 64
//   DHGR:  7 shared steps + 4 + 16 + 64                          =  91 steps / 64
// instead of 18 (10) steps per entry.  A thread owns one j-tile (its strings are
// decoded once into registers) and walks a chunk of i-tiles whose decoded strings
// the block stages in shared memory; a warp's 32 j-tiles make each row store
// 512 contiguous bytes.  A step is: LDS.U8 of S[a][b] (a is warp-uniform, so the
// 16-byte row is one broadcast wavefront), add, compare of the swap keys, min.
constexpr int kTreeThreads = 128;
constexpr int kTilesPerChunk = 16;   // i-tiles (of 8 rows) per block

template <int MODE>
struct Tree {
  static constexpr int kLeaf = MODE == IIV_MODE_HGR ? 4 : 3;
  static constexpr int kSfx = Mode<MODE>::kDots - kLeaf;      // shared pixels
  static constexpr int kSfxPad = (kSfx + 1) & ~1;
  // words of decoded i-side strings per tile: {lut row address, swap key} per pixel
  static constexpr int kTileWords = 2 * kSfxPad + 8 * 2 * 4;
};

constexpr uint32_t kInf = 0x3fffffffu;

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// One chain step for the two innermost pixels (0 and 1), which carry 85 % of the
// steps.  row = shared address of S[a_t][0] (16 bytes: a is warp-uniform, lanes
// differ in b only, so the load is one conflict-free broadcast wavefront);
// kf = a_t | a_{t+1} << 4; pb = b_t; kr = b_{t+1} | b_t << 4; f1 = F_{t+1};
// h2 = F_{t+2} + 1.
__device__ __forceinline__ uint32_t tree_step(uint32_t f1, uint32_t h2, uint32_t row,
                                              uint32_t kf, uint32_t pb, uint32_t kr) {
  uint32_t c = f1 + lds_u8(row + pb);
  if (kf == kr) c = min(c, h2);
  return c;
}

// The same step for the outer pixels, where registers matter more than bank
// conflicts: one per-lane register kr serves as swap key AND as index into the
// widened table S2[a][kr] = S[a][kr >> 4] (row2 = shared address of S2[a_t][0]).
__device__ __forceinline__ uint32_t tree_step2(uint32_t f1, uint32_t h2, uint32_t row2,
                                               uint32_t kf, uint32_t kr) {
  uint32_t c = f1 + lds_u8(row2 + kr);
  if (kf == kr) c = min(c, h2);
  return c;
}

template <int MODE, bool TRI, bool MULTI>
__global__ void __launch_bounds__(kTreeThreads, 4)
tree_kernel(const __grid_constant__ Lut lut, const __grid_constant__ Dests dests,
            uint32_t row_begin, uint32_t row_end) {
  using M = Mode<MODE>;
  using T = Tree<MODE>;
  constexpr int n = M::kDots, L = T::kLeaf;
  __shared__ __align__(16) uint8_t S[256];
  __shared__ __align__(16) uint8_t S2[16 * 256];
  __shared__ __align__(16) uint32_t idesc[kTilesPerChunk][T::kTileWords];

  const int tid = threadIdx.x;
  const int o = blockIdx.z;
  for (int k = tid; k < 256; k += kTreeThreads) S[k] = lut.s[k];
  for (int k = tid; k < 16 * 256; k += kTreeThreads)
    S2[k] = lut.s[(k >> 8) * 16 + ((k & 255) >> 4)];
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(S);
  const uint32_t s2_base = (uint32_t)__cvta_generic_to_shared(S2);

  // ---- i side: decode the chunk's strings into shared memory -----------------------
  const uint32_t tile0 = (row_begin >> 3) + blockIdx.y * kTilesPerChunk;
  const uint32_t tile_end = (row_end + 7) >> 3;
  constexpr int kItems = T::kSfx + 8 * L;            // (pixel) items per tile
  for (int item = tid; item < kTilesPerChunk * kItems; item += kTreeThreads) {
    const int tl = item / kItems, k = item - tl * kItems;
    const uint32_t tile = tile0 + tl;
    if (tile >= tile_end) continue;
    int v, t, slot;
    if (k < T::kSfx) {
      v = 0; t = L + k; slot = 2 * k;
    } else {
      v = (k - T::kSfx) / L; t = (k - T::kSfx) - v * L; slot = 2 * T::kSfxPad + 8 * v + 2 * t;
    }
    uint64_t lo; uint32_t hi;
    load_pixels<MODE>(o, tile * 8 + v, lo, hi);
    const uint32_t a = pixel_at(lo, hi, t);
    const uint32_t a1 = t + 1 < n ? pixel_at(lo, hi, t + 1) : 0xffffu;
    idesc[tl][slot] = t < 2 ? s_base + a * 16 : s2_base + a * 256;
    idesc[tl][slot + 1] = a | (a1 << 4);
  }

  // ---- j side: this thread's 8 strings, decoded into registers ----------------------
  const uint32_t jb = (blockIdx.x * kTreeThreads + tid) * 8;
  uint32_t sfx_kr[T::kSfx];
  uint32_t leaf_pb[8][2], leaf_kr[8][L];
  {
    uint64_t lo; uint32_t hi;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      load_pixels<MODE>(o, jb + v, lo, hi);
#pragma unroll
      for (int t = 0; t < L; ++t) {
        if (t < 2) leaf_pb[v][t] = pixel_at(lo, hi, t);
        leaf_kr[v][t] = pixel_at(lo, hi, t + 1) | (pixel_at(lo, hi, t) << 4);
      }
      if (v == 0) {
#pragma unroll
        for (int k = 0; k < T::kSfx; ++k) {
          const int t = L + k;
          sfx_kr[k] = (t + 1 < n ? pixel_at(lo, hi, t + 1) : 0u) | (pixel_at(lo, hi, t) << 4);
        }
      }
    }
  }
  __syncthreads();

  const size_t obase = ((size_t)o << (2 * M::kBits)) + jb;
  for (int tl = 0; tl < kTilesPerChunk; ++tl) {
    const uint32_t tile = tile0 + tl;
    if (tile >= tile_end) break;
    const uint32_t ib = tile * 8;
    const uint32_t* d = idesc[tl];

    // One finished row of the tile: mask (reference file layout keeps j < i only;
    // tiles are 8-aligned on both axes) and store 16 bytes per destination.
    auto emit = [&](int iv, const uint32_t (&r)[8]) {
      const uint32_t i = ib + iv;
      if (i < row_begin || i >= row_end) return;
      uint4 v = make_uint4(r[0] | (r[1] << 16), r[2] | (r[3] << 16), r[4] | (r[5] << 16),
                           r[6] | (r[7] << 16));
      if (TRI && jb >= ib) {
        if (jb > ib) {
          v = make_uint4(0, 0, 0, 0);
        } else {
          v.x &= (0 < iv ? 0xffffu : 0u) | (1 < iv ? 0xffff0000u : 0u);
          v.y &= (2 < iv ? 0xffffu : 0u) | (3 < iv ? 0xffff0000u : 0u);
          v.z &= (4 < iv ? 0xffffu : 0u) | (5 < iv ? 0xffff0000u : 0u);
          v.w &= (6 < iv ? 0xffffu : 0u) | (7 < iv ? 0xffff0000u : 0u);
        }
      }
      const size_t at = obase + ((size_t)i << M::kBits);
      if (MULTI) {
        if (dests.multicast) {
          multimem_st_v4(dests.p[0] + at, v);
        } else {
          for (int dd = 0; dd < dests.n; ++dd)
            *reinterpret_cast<uint4*>(dests.p[dd] + at) = v;
        }
      } else {
        *reinterpret_cast<uint4*>(dests.p[0] + at) = v;
      }
    };

    // triangular layout: the whole warp lies above the diagonal -> zeros only
    if (TRI && __all_sync(0xffffffffu, jb >= ib + 8)) {
      const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) emit(iv, z);
      continue;
    }

    // shared suffix: pixels n-1 .. L
    uint32_t f1, f2;    // F_{t+1}, F_{t+2}
    {
      const uint2 a = *reinterpret_cast<const uint2*>(d + 2 * (T::kSfx - 1));
      f1 = lds_u8(a.x + sfx_kr[T::kSfx - 1]);   // last pixel: no swap partner
      f2 = 0;
    }
#pragma unroll
    for (int k = T::kSfx - 2; k >= 0; --k) {
      const uint2 a = *reinterpret_cast<const uint2*>(d + 2 * k);
      const uint32_t f = tree_step2(f1, f2 + 1, a.x, a.y, sfx_kr[k]);
      f2 = f1;
      f1 = f;
    }
    const uint32_t* leaf = d + 2 * T::kSfxPad;     // [variant][pixel]{row, kf}
    if (MODE == IIV_MODE_HGR) {
      // pixels 3, 2 depend on bit 1 only
      uint32_t F2v[2][2], H3v[2][2], H2v[2][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const uint4 a23 = *reinterpret_cast<const uint4*>(leaf + 8 * (2 * mi) + 4);
#pragma unroll
        for (int mj = 0; mj < 2; ++mj) {
          const uint32_t f3 = tree_step2(f1, f2 + 1, a23.z, a23.w, leaf_kr[2 * mj][3]);
          F2v[mi][mj] = tree_step2(f3, f1 + 1, a23.x, a23.y, leaf_kr[2 * mj][2]);
          H3v[mi][mj] = f3 + 1;
          H2v[mi][mj] = F2v[mi][mj] + 1;
        }
      }
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) {
        const int mi = (iv >> 1) & 1;
        const uint4 a01 = *reinterpret_cast<const uint4*>(leaf + 8 * iv);
        uint32_t r[8];
#pragma unroll
        for (int jv = 0; jv < 8; ++jv) {
          const int mj = (jv >> 1) & 1;
          const uint32_t g1 = tree_step(F2v[mi][mj], H3v[mi][mj], a01.z, a01.w,
                                        leaf_pb[jv][1], leaf_kr[jv][1]);
          r[jv] = tree_step(g1, H2v[mi][mj], a01.x, a01.y, leaf_pb[jv][0], leaf_kr[jv][0]);
        }
        emit(iv, r);
      }
    } else {
      // pixel 2 <- bit 2; pixel 1 <- bits 1,2; pixel 0 <- bits 0,1,2
      uint32_t F2v[2][2], H2v[2][2];
#pragma unroll
      for (int i2 = 0; i2 < 2; ++i2) {
        const uint2 a2 = *reinterpret_cast<const uint2*>(leaf + 8 * (4 * i2) + 4);
#pragma unroll
        for (int j2 = 0; j2 < 2; ++j2) {
          F2v[i2][j2] = tree_step2(f1, f2 + 1, a2.x, a2.y, leaf_kr[4 * j2][2]);
          H2v[i2][j2] = F2v[i2][j2] + 1;
        }
      }
      const uint32_t h3 = f1 + 1;
#pragma unroll
      for (int i12 = 0; i12 < 4; ++i12) {
        const uint2 a1 = *reinterpret_cast<const uint2*>(leaf + 8 * (2 * i12) + 2);
        uint32_t F1v[4];
#pragma unroll
        for (int j12 = 0; j12 < 4; ++j12)
          F1v[j12] = tree_step(F2v[i12 >> 1][j12 >> 1], h3, a1.x, a1.y,
                               leaf_pb[2 * j12][1], leaf_kr[2 * j12][1]);
#pragma unroll
        for (int i0 = 0; i0 < 2; ++i0) {
          const int iv = 2 * i12 + i0;
          const uint2 a0 = *reinterpret_cast<const uint2*>(leaf + 8 * iv);
          uint32_t r[8];
#pragma unroll
          for (int jv = 0; jv < 8; ++jv)
            r[jv] = tree_step(F1v[jv >> 1], H2v[i12 >> 1][jv >> 2], a0.x, a0.y,
                              leaf_pb[jv][0], leaf_kr[jv][0]);
          emit(iv, r);
        }
      }
    }
  }
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
  if (algo == IIV_ALGO_CHAIN) {
    const uint32_t rows = row_end - row_begin;
    // gridDim.y <= 65535 holds: rows <= 16384.
    dim3 grid(N / (kChainThreads * kChainPerThread), rows, M::kOffsets);
    chain_kernel<MODE><<<grid, kChainThreads, 0, st>>>(
        lut, dests, row_begin, layout == IIV_LAYOUT_TRIANGULAR);
    IIV_LAUNCH_CHECK("chain_kernel");
    return 0;
  }
  const uint32_t tiles = ((row_end + 7) >> 3) - (row_begin >> 3);
  dim3 grid(N / (kTreeThreads * 8), (tiles + kTilesPerChunk - 1) / kTilesPerChunk,
            M::kOffsets);
  const bool tri = layout == IIV_LAYOUT_TRIANGULAR, multi = dests.n > 1 || dests.multicast;
  if (tri && multi)
    tree_kernel<MODE, true, true><<<grid, kTreeThreads, 0, st>>>(lut, dests, row_begin, row_end);
  else if (tri)
    tree_kernel<MODE, true, false><<<grid, kTreeThreads, 0, st>>>(lut, dests, row_begin, row_end);
  else if (multi)
    tree_kernel<MODE, false, true><<<grid, kTreeThreads, 0, st>>>(lut, dests, row_begin, row_end);
  else
    tree_kernel<MODE, false, false><<<grid, kTreeThreads, 0, st>>>(lut, dests, row_begin, row_end);
  IIV_LAUNCH_CHECK("tree_kernel");
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
  Dests dests = {};
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
  Dests dests = {};
  if (d_multicast_table) {
    dests.n = 1;
    dests.multicast = 1;
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

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
#include <mutex>

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
  uint32_t one;   // the constant 1, opaque to the compiler (see tree_step_x2)
};

// 16-byte store through a multicast (multimem) address: the switch replicates it
// into every member GPU's copy of the buffer.
__device__ __forceinline__ void multimem_st_v4(uint16_t* addr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr),
               "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
               "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}

// Programmatic dependent launch (sm_90+): a kernel launched with pdl_launch may start while
// its predecessor in the stream still runs -- it waits in grid_dependency_wait() before it
// touches anything the predecessor writes, and a predecessor calls launch_dependents() as
// soon as letting the successor's blocks queue up cannot hurt it.  Launched the ordinary
// way both calls are no-ops.
__device__ __forceinline__ void grid_dependency_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;");
}

template <typename... KArgs, typename... Args>
cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st,
                       Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <int MODE>
__global__ void pixel_prologue() {
  using M = Mode<MODE>;
  launch_dependents();
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
#ifndef IIV_TREE_MIN_BLOCKS
#define IIV_TREE_MIN_BLOCKS 5
#endif
constexpr int kTreeThreads = 128;
constexpr int kTilesPerChunk = 16;   // i-tiles (of 8 rows) per block

// Word offsets inside one i-tile descriptor in shared memory.  Scalar steps use
// {address of S2[a_t][0], swap key}; the two innermost pixels are walked for two
// rows at once in 16-bit halves and use {address of SP[a_t(row), a_t(row')][0],
// swap keys of both rows}.
template <int MODE>
struct Tree {
  static constexpr bool kHgr = MODE == IIV_MODE_HGR;
  static constexpr int kLeaf = kHgr ? 4 : 3;
  static constexpr int kSfx = Mode<MODE>::kDots - kLeaf;      // shared pixels
  static constexpr int kSfxPad = (kSfx + 1) & ~1;
  static constexpr int kMidOff = 2 * kSfxPad;   // HGR (mi, pixel 2|3); DHGR (i2, pixel 2)
  static constexpr int kMidWords = kHgr ? 8 : 4;
  static constexpr int kP1Off = kMidOff + kMidWords;   // pixel 1, row pairs
  static constexpr int kP1Pairs = kHgr ? 4 : 2;
  static constexpr int kP0Off = kP1Off + 2 * kP1Pairs;   // pixel 0, 4 row pairs
  static constexpr int kTileWords = kP0Off + 8;
  static constexpr int kItems = kSfx + kMidWords / 2 + kP1Pairs + 4;
};

// One chain step (scalar).  s = S2[a_t][kr] where S2[a][k] = S[a][k >> 4]; kf = a_t |
// a_{t+1} << 4; kr = b_{t+1} | b_t << 4 serves as swap key AND table index; f1 =
// F_{t+1}; h2 = F_{t+2} + 1.
__device__ __forceinline__ uint32_t tree_step(uint32_t f1, uint32_t h2, uint32_t s,
                                              uint32_t kf, uint32_t kr) {
  uint32_t c = f1 + s;
  if (kf == kr) c = min(c, h2);
  return c;
}

// The same step for two rows at once, values in 16-bit halves (all < 0x8000).
// s2 = SP[a_t, a_t'][b_t], SP[a, a'][b] = S[a][b] | S[a'][b] << 16 (a, a' are
// warp-uniform and lanes differ in b only: one conflict-free broadcast wavefront);
// kf2 / kr2 = swap keys of both rows / of b in both halves.  A half whose keys differ
// gets bit 15 set in the swap alternative, which the unsigned min then never takes:
// VIADDMNMX.U16x2 does add + min for both rows.
//
// `one` is the constant 1 in a register the compiler cannot see through: the adds
// written as x * one + y become IMAD on the FMA pipe, which otherwise idles while the
// ALU pipe (LOP3 / VIADDMNMX / PRMT, half rate) is the bottleneck.
__device__ __forceinline__ uint32_t tree_step_x2(uint32_t f1, uint32_t h2, uint32_t s2,
                                                 uint32_t kf2, uint32_t kr2, uint32_t one) {
  const uint32_t x = kf2 ^ kr2;
  const uint32_t alt = ((x * one + 0x7fff7fffu) & 0x80008000u) | h2;
  return __viaddmin_u16x2(f1, s2, alt);
}

template <int MODE, bool TRI, bool MULTI>
__global__ void __launch_bounds__(kTreeThreads, IIV_TREE_MIN_BLOCKS)
tree_kernel(const __grid_constant__ Lut lut, const __grid_constant__ Dests dests,
            uint32_t row_begin, uint32_t row_end) {
  using M = Mode<MODE>;
  using T = Tree<MODE>;
  constexpr int n = M::kDots, L = T::kLeaf;
  __shared__ __align__(16) uint8_t S2[16 * 256];
  __shared__ __align__(16) uint32_t SP[256 * 16];
  __shared__ __align__(16) uint32_t idesc[kTilesPerChunk][T::kTileWords];

  const int tid = threadIdx.x;
  const int o = blockIdx.z;
  for (int k = tid; k < 16 * 256; k += kTreeThreads) {
    S2[k] = lut.s[(k >> 8) * 16 + ((k & 255) >> 4)];
    // k = (a << 4 | a') << 4 | b
    SP[k] = (uint32_t)lut.s[(k >> 8) * 16 + (k & 15)] |
            ((uint32_t)lut.s[((k >> 4) & 15) * 16 + (k & 15)] << 16);
  }
  const uint32_t one = dests.one;
  // descriptors hold byte offsets into S2 / SP; lanes add their own b-dependent part
  const auto s2_at = [&](uint32_t off) -> uint32_t { return S2[off]; };
  const auto sp_at = [&](uint32_t off) -> uint32_t {
    return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(SP) + off);
  };

  // ---- i side: decode the chunk's strings into shared memory -----------------------
  const uint32_t tile0 = (row_begin >> 3) + blockIdx.y * kTilesPerChunk;
  const uint32_t tile_end = (row_end + 7) >> 3;
  for (int item = tid; item < kTilesPerChunk * T::kItems; item += kTreeThreads) {
    const int tl = item / T::kItems;
    int k = item - tl * T::kItems;
    const uint32_t tile = tile0 + tl;
    if (tile >= tile_end) continue;
    // scalar item: (value v, pixel t); paired item: (v, v2, t)
    int v, v2 = -1, t, slot;
    if (k < T::kSfx) {
      v = 0; t = L + k; slot = 2 * k;
    } else if ((k -= T::kSfx) < T::kMidWords / 2) {
      if (T::kHgr) { v = 2 * (k >> 1); t = 2 + (k & 1); }   // (mi, pixel 2 | 3)
      else { v = 4 * k; t = 2; }                              // (i2, pixel 2)
      slot = T::kMidOff + 2 * k;
    } else if ((k -= T::kMidWords / 2) < T::kP1Pairs) {
      if (T::kHgr) { v = 2 * k; v2 = 2 * k + 1; }            // rows 2q, 2q+1
      else { v = 4 * k; v2 = 4 * k + 2; }                     // i12 = 2q, 2q+1
      t = 1; slot = T::kP1Off + 2 * k;
    } else {
      k -= T::kP1Pairs;
      v = 2 * k; v2 = 2 * k + 1; t = 0; slot = T::kP0Off + 2 * k;
    }
    uint64_t lo; uint32_t hi;
    load_pixels<MODE>(o, tile * 8 + v, lo, hi);
    const uint32_t a = pixel_at(lo, hi, t);
    const uint32_t kf = a | ((t + 1 < n ? pixel_at(lo, hi, t + 1) : 0xffu) << 4);
    if (v2 < 0) {
      idesc[tl][slot] = a * 256;
      idesc[tl][slot + 1] = kf;
    } else {
      load_pixels<MODE>(o, tile * 8 + v2, lo, hi);
      const uint32_t a2 = pixel_at(lo, hi, t);
      const uint32_t kf2 = a2 | (pixel_at(lo, hi, t + 1) << 4);
      idesc[tl][slot] = (a << 4 | a2) << 6;
      idesc[tl][slot + 1] = kf | (kf2 << 16);
    }
  }

  // ---- j side: this thread's 8 strings, decoded into registers ----------------------
  const uint32_t jb = (blockIdx.x * kTreeThreads + tid) * 8;
  uint32_t sfx_kr[T::kSfx];
  uint32_t leaf_pb4[8][2], leaf_kr2[8][2], mid_kr[8][L];
  {
    uint64_t lo; uint32_t hi;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      load_pixels<MODE>(o, jb + v, lo, hi);
#pragma unroll
      for (int t = 0; t < L; ++t) {
        const uint32_t kr = pixel_at(lo, hi, t + 1) | (pixel_at(lo, hi, t) << 4);
        if (t < 2) {
          leaf_pb4[v][t] = pixel_at(lo, hi, t) * 4;
          leaf_kr2[v][t] = kr * 0x10001u;
        }
        mid_kr[v][t] = kr;
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
    auto emit = [&](int iv, uint4 v) {
      const uint32_t i = ib + iv;
      if (i < row_begin || i >= row_end) return;
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
    // r[jv] holds rows (iv, iv + 1) in its halves: un-zip into the two rows
    auto emit_pair = [&](int iv, const uint32_t (&r)[8]) {
      emit(iv, make_uint4(__byte_perm(r[0], r[1], 0x5410), __byte_perm(r[2], r[3], 0x5410),
                          __byte_perm(r[4], r[5], 0x5410), __byte_perm(r[6], r[7], 0x5410)));
      emit(iv + 1, make_uint4(__byte_perm(r[0], r[1], 0x7632), __byte_perm(r[2], r[3], 0x7632),
                              __byte_perm(r[4], r[5], 0x7632), __byte_perm(r[6], r[7], 0x7632)));
    };

    // triangular layout: the whole warp lies above the diagonal -> zeros only
    if (TRI && __all_sync(0xffffffffu, jb >= ib + 8)) {
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) emit(iv, make_uint4(0, 0, 0, 0));
      continue;
    }

    // shared suffix: pixels n-1 .. L
    uint32_t f1, f2;    // F_{t+1}, F_{t+2}
    {
      const uint2 a = *reinterpret_cast<const uint2*>(d + 2 * (T::kSfx - 1));
      f1 = s2_at(a.x + sfx_kr[T::kSfx - 1]);   // last pixel: no swap partner
      f2 = 0;
    }
#pragma unroll
    for (int k = T::kSfx - 2; k >= 0; --k) {
      const uint2 a = *reinterpret_cast<const uint2*>(d + 2 * k);
      const uint32_t f = tree_step(f1, f2 + 1, s2_at(a.x + sfx_kr[k]), a.y, sfx_kr[k]);
      f2 = f1;
      f1 = f;
    }
    if (T::kHgr) {
      // pixels 3, 2 depend on bit 1 only: 4 scalar (mi, mj) combinations
      uint32_t F2p[2][2], H3p[2][2], H2p[2][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const uint4 a23 = *reinterpret_cast<const uint4*>(d + T::kMidOff + 4 * mi);
#pragma unroll
        for (int mj = 0; mj < 2; ++mj) {
          const uint32_t f3 = tree_step(f1, f2 + 1, s2_at(a23.z + mid_kr[2 * mj][3]), a23.w,
                                        mid_kr[2 * mj][3]);
          const uint32_t g2 = tree_step(f3, f1 + 1, s2_at(a23.x + mid_kr[2 * mj][2]), a23.y,
                                        mid_kr[2 * mj][2]);
          F2p[mi][mj] = g2 * 0x10001u;
          H3p[mi][mj] = (f3 + 1) * 0x10001u;
          H2p[mi][mj] = (g2 + 1) * 0x10001u;
        }
      }
      // pixels 1, 0: rows (2q, 2q+1) together, both have bit 1 = q & 1
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int mi = q & 1;
        const uint2 a1 = *reinterpret_cast<const uint2*>(d + T::kP1Off + 2 * q);
        const uint2 a0 = *reinterpret_cast<const uint2*>(d + T::kP0Off + 2 * q);
        uint32_t r[8];
#pragma unroll
        for (int jv = 0; jv < 8; ++jv) {
          const int mj = (jv >> 1) & 1;
          const uint32_t g1 = tree_step_x2(F2p[mi][mj], H3p[mi][mj],
                                           sp_at(leaf_pb4[jv][1] * one + a1.x), a1.y,
                                           leaf_kr2[jv][1], one);
          r[jv] = tree_step_x2(g1, H2p[mi][mj], sp_at(leaf_pb4[jv][0] * one + a0.x), a0.y,
                               leaf_kr2[jv][0], one);
        }
        emit_pair(2 * q, r);
      }
    } else {
      // pixel 2 <- bit 2; pixel 1 <- bits 1,2; pixel 0 <- bits 0,1,2
      uint32_t F2p[2][2], H2p[2][2];
#pragma unroll
      for (int i2 = 0; i2 < 2; ++i2) {
        const uint2 a2 = *reinterpret_cast<const uint2*>(d + T::kMidOff + 2 * i2);
#pragma unroll
        for (int j2 = 0; j2 < 2; ++j2) {
          const uint32_t g2 = tree_step(f1, f2 + 1, s2_at(a2.x + mid_kr[4 * j2][2]), a2.y,
                                        mid_kr[4 * j2][2]);
          F2p[i2][j2] = g2 * 0x10001u;
          H2p[i2][j2] = (g2 + 1) * 0x10001u;
        }
      }
      const uint32_t h3p = (f1 + 1) * 0x10001u;
#pragma unroll
      for (int q1 = 0; q1 < 2; ++q1) {
        // pixel 1 for i12 = 2 q1 and 2 q1 + 1 (both have bit 2 = q1)
        const uint2 a1 = *reinterpret_cast<const uint2*>(d + T::kP1Off + 2 * q1);
        uint32_t F1x[4];
#pragma unroll
        for (int j12 = 0; j12 < 4; ++j12)
          F1x[j12] = tree_step_x2(F2p[q1][j12 >> 1], h3p,
                                  sp_at(leaf_pb4[2 * j12][1] * one + a1.x), a1.y,
                                  leaf_kr2[2 * j12][1], one);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int q0 = 2 * q1 + half;           // rows 2 q0, 2 q0 + 1: i12 = q0
          const uint2 a0 = *reinterpret_cast<const uint2*>(d + T::kP0Off + 2 * q0);
          uint32_t r[8];
#pragma unroll
          for (int jv = 0; jv < 8; ++jv) {
            const uint32_t parent = __byte_perm(F1x[jv >> 1], 0, half ? 0x3232 : 0x1010);
            r[jv] = tree_step_x2(parent, H2p[q1][jv >> 2], sp_at(leaf_pb4[jv][0] * one + a0.x),
                                 a0.y, leaf_kr2[jv][0], one);
          }
          emit_pair(2 * q0, r);
        }
      }
    }
  }
}

// ---- ALGO_SPLIT: the chain cut in two, each half tabulated -------------------------
// The recurrence is a product of 2x2 (min,+) matrices, one per pixel, so it can be cut
// at pixel c:
//     entry = min(D_c + F_c,  D_{c-1} + 1 + F_{c+1} if pixels (c-1, c) are a swapped pair)
// with D the forward form over pixels 0..c-1 (chain_kernel) and F the backward form over
// pixels c..n-1 (tree_kernel).  Pixel t is a rotation of dots t..t+3 (colours.py:100-134),
// and a dot is fed by few bits of the masked value (screen.py:741-789), so each half
// depends on a WINDOW of the value's bits only:
//     HGR  (c = 8): pixels 0..8 <- 9 bits (header, palette bit, low body), pixels 8..17 <- 9
//     DHGR (c = 4): pixels 0..4 <- bits 0..7,                       pixels 4..9  <- bits 4..12
// (tests/test_oracle_tables.py::test_split_windows checks the windows against the oracle's
// pixel strings).  split_prologue tabulates A = (D_c, D_{c-1}+1 | INF) over pairs of A
// windows and B = (F_c, F_{c+1}) over pairs of B windows -- 2^16..2^18 pairs per offset
// instead of 2^26 / 2^28 -- and an entry of the big table is then one packed add and one
// VIADDMNMX.U16x2 for two entries: the generator is a write stream.
//
// A thread owns 8 consecutive j inside the A window (their 8 A pairs stay in registers)
// and walks the 32 values of the 5 bits of j outside that window; every step is one
// 4-byte load of B (L1), two byte permutes, 8 packed ALU ops and one 16-byte store; a
// warp's 32 threads cover 256 consecutive j, so every store instruction writes 512
// contiguous bytes.
#ifndef IIV_SPLIT_THREADS
#define IIV_SPLIT_THREADS 256
#endif
#ifndef IIV_SPLIT_EPT
#define IIV_SPLIT_EPT 4
#endif
#ifndef IIV_SPLIT_MIN_BLOCKS
#define IIV_SPLIT_MIN_BLOCKS 8
#endif
constexpr int kSplitThreads = IIV_SPLIT_THREADS;
constexpr int kSplitEpt = IIV_SPLIT_EPT;   // values of the outside bits one thread walks
constexpr uint32_t kSplitInf = 0x4000;   // + any F stays below 0x8000 (n * 255 <= 4590)

template <uint32_t MASK>
using Bits = Ext<MASK>;   // iiv_common.cuh

// Windows per (mode, window set).  HGR has one set per offset (the palette bit that shifts
// the body's dots is bit 10 at offset 0 and bit 3 at offset 1); DHGR's dots are the value.
// kLateA: two bits of A's window that reach no pixel before kSharedA; kEarlyB: two bits of
// B's window that reach no pixel from kSharedB on -- the prologue walks the part of a chain
// those bits cannot change once for their four values.
template <int MODE, int WIN>
struct SplitWin;
template <>
struct SplitWin<IIV_MODE_HGR, 0> {
  static constexpr int kCut = 8, kSharedA = 6, kSharedB = 12;
  static constexpr uint32_t kMaskA = 0x04ff, kMaskB = 0x3fe0, kLateA = 0x00c0, kEarlyB = 0x0060;
};
template <>
struct SplitWin<IIV_MODE_HGR, 1> {
  static constexpr int kCut = 8, kSharedA = 6, kSharedB = 12;
  static constexpr uint32_t kMaskA = 0x01ff, kMaskB = 0x3fc8, kLateA = 0x0180, kEarlyB = 0x00c0;
};
template <>
struct SplitWin<IIV_MODE_DHGR, 0> {
  static constexpr int kCut = 4, kSharedA = 3, kSharedB = 6;
  static constexpr uint32_t kMaskA = 0x00ff, kMaskB = 0x1ff0, kLateA = 0x00c0, kEarlyB = 0x0030;
};
template <int MODE>
struct SplitDims {
  static constexpr int kWins = MODE == IIV_MODE_HGR ? 2 : 1;   // window sets
  static constexpr int kA = Bits<SplitWin<MODE, 0>::kMaskA>::count;
  static constexpr int kB = Bits<SplitWin<MODE, 0>::kMaskB>::count;
  static constexpr size_t kWordsA = (size_t)Mode<MODE>::kOffsets << (2 * kA);   // uint16 pairs
  static constexpr size_t kWordsB = (size_t)Mode<MODE>::kOffsets << (2 * kB);   // uint32
};

// A[o][xi][0][xj] = D_c, A[o][xi][1][xj] = D_{c-1} + 1 or INF (uint16);
// B[o][yi][yj] = F_c | F_{c+1} << 16.  A thread takes one xi (yi) and the four xj (yj) that
// differ in the late (early) bits.
// The pixel strings of one value computed in registers (what pixel_prologue stores for the
// chain and tree kernels): the split prologue needs five strings per thread and is bound by
// latency, not issue slots -- deriving them costs less than the loads did, and the
// pixel_prologue launch is gone (HGR step -2.5 %, DHGR -7 %).
template <int MODE>
__device__ __forceinline__ void make_pixels(int o, uint32_t v, uint64_t& lo, uint32_t& hi) {
  using M = Mode<MODE>;
  const uint32_t dots = to_dots<MODE>(v, o);
  lo = 0;
  hi = 0;
#pragma unroll
  for (int t = 0; t < M::kDots; ++t) {
    const uint64_t p = nominal_pixel(dots, t, M::phase(o));
    if (t < 16)
      lo |= p << (4 * t);
    else
      hi |= (uint32_t)p << (4 * (t - 16));
  }
}

template <int MODE, int WIN>
__device__ __forceinline__ void split_tabulate(const uint8_t* S, int o, int which, uint32_t idx,
                                               uint16_t* ta, uint32_t* tb) {
  using M = Mode<MODE>;
  using W = SplitWin<MODE, WIN>;
  using BA = Bits<W::kMaskA>;
  using BB = Bits<W::kMaskB>;
  constexpr int n = M::kDots, c = W::kCut;
  uint64_t alo, blo[4];
  uint32_t ahi, bhi[4];
  if (which == 0) {
    constexpr uint32_t NA = 1u << BA::count, late = BA::ext(W::kLateA);
    using BL = Bits<late>;
    using BR = Bits<~late & (NA - 1u)>;
    static_assert(BL::count == 2, "two late bits");
    if (idx >= NA * NA / 4) return;
    const uint32_t xi = idx / (NA / 4), xj0 = BR::dep(idx % (NA / 4));
    make_pixels<MODE>(o, BA::dep(xi), alo, ahi);
#pragma unroll
    for (int v = 0; v < 4; ++v) make_pixels<MODE>(o, BA::dep(xj0 | BL::dep(v)), blo[v], bhi[v]);
    uint32_t d2 = 0, d1 = 0, pa = 0, pb = 0;
#pragma unroll
    for (int t = 0; t < W::kSharedA; ++t) {
      const uint32_t a = pixel_at(alo, ahi, t), b = pixel_at(blo[0], bhi[0], t);
      uint32_t cur = d1 + S[a * 16 + b];
      if (t >= 1 && pa == b && a == pb) cur = min(cur, d2 + 1u);
      d2 = d1; d1 = cur; pa = a; pb = b;
    }
    uint16_t* row = ta + (((size_t)o * NA + xi) * 2) * NA + xj0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      uint32_t e2 = d2, e1 = d1, qa = pa, qb = pb;
#pragma unroll
      for (int t = W::kSharedA; t < c; ++t) {
        const uint32_t a = pixel_at(alo, ahi, t), b = pixel_at(blo[v], bhi[v], t);
        uint32_t cur = e1 + S[a * 16 + b];
        if (t >= 1 && qa == b && a == qb) cur = min(cur, e2 + 1u);
        e2 = e1; e1 = cur; qa = a; qb = b;
      }
      const uint32_t a = pixel_at(alo, ahi, c), b = pixel_at(blo[v], bhi[v], c);
      row[BL::dep(v)] = (uint16_t)e1;
      row[NA + BL::dep(v)] = (uint16_t)((qa == b && a == qb) ? e2 + 1u : kSplitInf);
    }
  } else {
    constexpr uint32_t NB = 1u << BB::count, early = BB::ext(W::kEarlyB);
    using BE = Bits<early>;
    using BR = Bits<~early & (NB - 1u)>;
    static_assert(BE::count == 2, "two early bits");
    if (idx >= NB * NB / 4) return;
    const uint32_t yi = idx / (NB / 4), yj0 = BR::dep(idx % (NB / 4));
    make_pixels<MODE>(o, BB::dep(yi), alo, ahi);
#pragma unroll
    for (int v = 0; v < 4; ++v) make_pixels<MODE>(o, BB::dep(yj0 | BE::dep(v)), blo[v], bhi[v]);
    uint32_t f1 = 0, f2 = 0, na = 0, nb = 0;   // F_{t+1}, F_{t+2}, pixels t+1
#pragma unroll
    for (int t = n - 1; t >= W::kSharedB; --t) {
      const uint32_t a = pixel_at(alo, ahi, t), b = pixel_at(blo[0], bhi[0], t);
      uint32_t cur = f1 + S[a * 16 + b];
      if (t + 1 < n && a == nb && na == b) cur = min(cur, f2 + 1u);
      f2 = f1; f1 = cur; na = a; nb = b;
    }
    uint32_t* row = tb + ((size_t)o * NB + yi) * NB + yj0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      uint32_t g1 = f1, g2 = f2, ma = na, mb = nb;
#pragma unroll
      for (int t = W::kSharedB - 1; t >= c; --t) {
        const uint32_t a = pixel_at(alo, ahi, t), b = pixel_at(blo[v], bhi[v], t);
        uint32_t cur = g1 + S[a * 16 + b];
        if (t + 1 < n && a == mb && ma == b) cur = min(cur, g2 + 1u);
        g2 = g1; g1 = cur; ma = a; mb = b;
      }
      row[BE::dep(v)] = g1 | (g2 << 16);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256)
split_prologue(const __grid_constant__ Lut lut, uint16_t* __restrict__ ta,
               uint32_t* __restrict__ tb) {
  __shared__ uint8_t S[256];
  launch_dependents();
  S[threadIdx.x] = lut.s[threadIdx.x];
  __syncthreads();
  const int o = blockIdx.z, which = blockIdx.y;
  const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
  if (MODE == IIV_MODE_HGR && o == 1)
    split_tabulate<MODE, SplitDims<MODE>::kWins - 1>(S, o, which, idx, ta, tb);
  else
    split_tabulate<MODE, 0>(S, o, which, idx, ta, tb);
}

template <int MODE, int WIN, bool TRI, bool MULTI>
__device__ __forceinline__ void split_rows(int o, const uint16_t* __restrict__ ta,
                                           const uint32_t* __restrict__ tb, const Dests& dests,
                                           uint32_t row_begin, uint32_t row_end) {
  using M = Mode<MODE>;
  using W = SplitWin<MODE, WIN>;
  using BA = Bits<W::kMaskA>;
  using BB = Bits<W::kMaskB>;
  using BO = Bits<~W::kMaskA & ((1u << M::kBits) - 1u)>;   // the bits of j outside A's window
  static_assert((W::kMaskA & 0xffu) == 0xffu, "a warp's 32 octets are 256 consecutive j");
  static_assert((W::kMaskB & 7u) == 0u, "B is constant over an octet of j");
  constexpr uint32_t NA = 1u << BA::count, NB = 1u << BB::count;
  constexpr int kThreadsPerRow = NA / 8, kRows = kSplitThreads / kThreadsPerRow;

  constexpr uint32_t kGroups = (1u << BO::count) / kSplitEpt;   // threads sharing (i, octet)
  static_assert(kGroups * kSplitEpt == (1u << BO::count), "EPT divides the outside values");
  const uint32_t e0 = (blockIdx.x % kGroups) * kSplitEpt;
  const uint32_t i = row_begin + (blockIdx.x / kGroups) * kRows + threadIdx.x / kThreadsPerRow;
  if (i >= row_end) return;
  const uint32_t xj = (threadIdx.x % kThreadsPerRow) * 8;   // A index of the octet's first j
  const uint32_t jb = BA::dep(xj);                          // that j, outside bits zero
  const uint16_t* arow = ta + (((size_t)o * NA + BA::ext(i)) * 2) * NA + xj;
  grid_dependency_wait();   // the A / B tables
  const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(arow));
  const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(arow + NA));
  const uint32_t J0 = BO::dep(e0);
  const uint32_t* brow = tb + ((size_t)o * NB + BB::ext(i)) * NB + BB::ext(jb | J0);
  const size_t at = ((size_t)o << (2 * M::kBits)) + ((size_t)i << M::kBits) + (jb | J0);
  const int below = (int)i - (int)(jb | J0);   // entries j < i of the octet at J: below - J

#pragma unroll
  for (uint32_t e = 0; e < (uint32_t)kSplitEpt; ++e) {
    const uint32_t J = BO::dep(e);   // compile-time after unrolling (e0's bits lie above)
    uint4 v;
    if (TRI && below - (int)J <= 0) {
      v = make_uint4(0, 0, 0, 0);
    } else {
      const uint32_t b = __ldg(brow + BB::ext(J));
      const uint32_t b0 = __byte_perm(b, 0, 0x1010), b1 = __byte_perm(b, 0, 0x3232);
      v.x = __viaddmin_u16x2(a1.x, b1, a0.x + b0);
      v.y = __viaddmin_u16x2(a1.y, b1, a0.y + b0);
      v.z = __viaddmin_u16x2(a1.z, b1, a0.z + b0);
      v.w = __viaddmin_u16x2(a1.w, b1, a0.w + b0);
      if (TRI && below - (int)J < 8) {
        const int d = below - (int)J;
        v.x &= (0 < d ? 0xffffu : 0u) | (1 < d ? 0xffff0000u : 0u);
        v.y &= (2 < d ? 0xffffu : 0u) | (3 < d ? 0xffff0000u : 0u);
        v.z &= (4 < d ? 0xffffu : 0u) | (5 < d ? 0xffff0000u : 0u);
        v.w &= (6 < d ? 0xffffu : 0u) | (7 < d ? 0xffff0000u : 0u);
      }
    }
    if (MULTI) {
      if (dests.multicast) {
        multimem_st_v4(dests.p[0] + at + J, v);
      } else {
        for (int dd = 0; dd < dests.n; ++dd)
          *reinterpret_cast<uint4*>(dests.p[dd] + at + J) = v;
      }
    } else {
      *reinterpret_cast<uint4*>(dests.p[0] + at + J) = v;
    }
  }
}

template <int MODE, bool TRI, bool MULTI>
__global__ void __launch_bounds__(kSplitThreads, IIV_SPLIT_MIN_BLOCKS)
split_kernel(const uint16_t* __restrict__ ta, const uint32_t* __restrict__ tb,
             const __grid_constant__ Dests dests, uint32_t row_begin, uint32_t row_end) {
  const int o = blockIdx.y;
  if (MODE == IIV_MODE_HGR && o == 1)
    split_rows<MODE, SplitDims<MODE>::kWins - 1, TRI, MULTI>(o, ta, tb, dests, row_begin, row_end);
  else
    split_rows<MODE, 0, TRI, MULTI>(o, ta, tb, dests, row_begin, row_end);
}

// Scratch for the A / B tables of one generate call: stream-ordered allocations from a pool
// the library keeps per device (release threshold unlimited, so a steady stream of calls
// allocates nothing).  The tables depend on the LUT, so they cannot live in __device__
// globals shared by concurrent calls.
int split_scratch(void** p, size_t bytes, cudaStream_t st) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  IIV_CUDA(cudaGetDevice(&dev));
  IIV_REQUIRE(dev >= 0 && dev < 64, "device %d out of range", dev);
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      IIV_CUDA(cudaMemPoolCreate(&pools[dev], &props));
      uint64_t keep = ~(uint64_t)0;
      IIV_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    pool = pools[dev];
  }
  IIV_CUDA(cudaMallocFromPoolAsync(p, bytes, pool, st));
  return 0;
}

template <int MODE>
int generate_split(const Lut& lut, const Dests& dests, uint32_t row_begin, uint32_t row_end,
                   bool tri, bool multi, cudaStream_t st) {
  using M = Mode<MODE>;
  using D = SplitDims<MODE>;
  constexpr size_t bytes_a = D::kWordsA * 2 * sizeof(uint16_t), bytes_b = D::kWordsB * 4;
  void* scratch = nullptr;
  const int rc = split_scratch(&scratch, bytes_a + bytes_b, st);
  if (rc) return rc;
  uint16_t* ta = reinterpret_cast<uint16_t*>(scratch);
  uint32_t* tb = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(scratch) + bytes_a);
  constexpr uint32_t pairs = 1u << (2 * (D::kA > D::kB ? D::kA : D::kB));
  constexpr int rows_per_block = kSplitThreads / ((1 << D::kA) / 8);
  constexpr int groups = (1 << (M::kBits - D::kA)) / kSplitEpt;
  const dim3 grid((row_end - row_begin + rows_per_block - 1) / rows_per_block * groups,
                  M::kOffsets);
  const dim3 block(kSplitThreads);
  // The generator is a programmatic dependent of the prologue (A / B tables -> generator):
  // its launch latency and its blocks' preamble hide behind the prologue.  (One prologue + generator pair per offset, so that offset o + 1's tables are
  // made while offset o's rows drain, measured 3-8 % SLOWER: the second prologue is
  // latency-bound and the generator cannot start before it ends.)
  // (The prologue itself is an ordinary launch: it overwrites scratch the previous call's
  // generator may still be reading.)
  split_prologue<MODE><<<dim3(pairs / 4 / 256, 2, M::kOffsets), 256, 0, st>>>(lut, ta, tb);
  cudaError_t launched = cudaGetLastError();
  if (launched == cudaSuccess) {
    if (tri && multi)
      launched = pdl_launch(split_kernel<MODE, true, true>, grid, block, st, ta, tb, dests, row_begin, row_end);
    else if (tri)
      launched = pdl_launch(split_kernel<MODE, true, false>, grid, block, st, ta, tb, dests, row_begin, row_end);
    else if (multi)
      launched = pdl_launch(split_kernel<MODE, false, true>, grid, block, st, ta, tb, dests, row_begin, row_end);
    else
      launched = pdl_launch(split_kernel<MODE, false, false>, grid, block, st, ta, tb, dests, row_begin, row_end);
  }
  const cudaError_t freed = cudaFreeAsync(scratch, st);
  if (launched != cudaSuccess) return cuda_fail(launched, "split_kernel");
  if (freed != cudaSuccess) return cuda_fail(freed, "cudaFreeAsync");
  return 0;
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
  const bool tri = layout == IIV_LAYOUT_TRIANGULAR, multi = dests.n > 1 || dests.multicast;
  // the split generator derives the strings it needs in registers (no pixel-string table)
  if (algo != IIV_ALGO_CHAIN && algo != IIV_ALGO_TREE)
    return generate_split<MODE>(lut, dests, row_begin, row_end, tri, multi, st);
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

// The bit windows ALGO_SPLIT cuts the masked value into (host-only query: the CPU tests
// check them against the oracle's pixel strings).
extern "C" int iiv_table_split_windows(int mode, int offset, int* cut, uint32_t* mask_a,
                                       uint32_t* mask_b) {
  IIV_REQUIRE(cut && mask_a && mask_b, "null pointer");
  if (mode == IIV_MODE_HGR && offset == 0) {
    using W = SplitWin<IIV_MODE_HGR, 0>;
    *cut = W::kCut; *mask_a = W::kMaskA; *mask_b = W::kMaskB;
  } else if (mode == IIV_MODE_HGR && offset == 1) {
    using W = SplitWin<IIV_MODE_HGR, 1>;
    *cut = W::kCut; *mask_a = W::kMaskA; *mask_b = W::kMaskB;
  } else if (mode == IIV_MODE_DHGR && offset >= 0 && offset < 4) {
    using W = SplitWin<IIV_MODE_DHGR, 0>;
    *cut = W::kCut; *mask_a = W::kMaskA; *mask_b = W::kMaskB;
  } else {
    IIV_REQUIRE(false, "bad mode %d / offset %d", mode, offset);
  }
  return 0;
}

extern "C" int iiv_table_generate(int mode, const int32_t* h_lut,
                                  uint16_t* d_table, uint32_t row_begin,
                                  uint32_t row_end, int layout, int algo,
                                  void* stream) {
  IIV_REQUIRE(h_lut && d_table, "null pointer");
  Dests dests = {};
  dests.n = 1;
  dests.one = 1;
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
  dests.one = 1;
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

// The device-to-host leg of compute_edit_distance.  In the reference's file layout only
// j < i can be nonzero (make_data_tables.py:156-172), so moving whole rows would spend half
// of the PCIe time on zeros: the rows go home in bands, each as one 2-D copy that is only as
// wide as the band's last row needs.  The rest of the destination is left alone -- the
// caller hands in a buffer whose other bytes are already zero.
extern "C" int iiv_table_download(int mode, const uint16_t* d_table, uint16_t* h_table,
                                  uint32_t row_begin, uint32_t row_end, int layout, int bands,
                                  void* stream) {
  IIV_REQUIRE(mode == IIV_MODE_HGR || mode == IIV_MODE_DHGR, "bad mode %d", mode);
  IIV_REQUIRE(d_table && h_table, "null pointer");
  IIV_REQUIRE(layout == IIV_LAYOUT_TRIANGULAR || layout == IIV_LAYOUT_SYMMETRIC,
              "bad layout %d", layout);
  const uint32_t bits = mode == IIV_MODE_HGR ? 14 : 13;
  const uint32_t n = 1u << bits, n_off = mode == IIV_MODE_HGR ? 2 : 4;
  IIV_REQUIRE(row_begin <= row_end && row_end <= n, "bad row range [%u,%u)", row_begin, row_end);
  IIV_REQUIRE(bands >= 1 && bands <= 4096, "bad band count %d", bands);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t pitch = (size_t)n * 2;
  if (row_begin == row_end) return 0;
  if (layout == IIV_LAYOUT_SYMMETRIC) bands = 1;
  const uint32_t rows = row_end - row_begin;
  const uint32_t per = (rows + (uint32_t)bands - 1) / (uint32_t)bands;
  for (uint32_t o = 0; o < n_off; ++o) {
    for (uint32_t r0 = row_begin; r0 < row_end; r0 += per) {
      const uint32_t r1 = r0 + per < row_end ? r0 + per : row_end;
      // columns j < r1 - 1 cover every j < i of the band; rounded up to 64 bytes
      size_t width = layout == IIV_LAYOUT_SYMMETRIC ? pitch : (((size_t)(r1 - 1) * 2 + 63) & ~(size_t)63);
      if (width > pitch) width = pitch;
      if (width == 0) continue;
      const size_t at = ((size_t)o << (2 * bits)) + ((size_t)r0 << bits);
      IIV_CUDA(cudaMemcpy2DAsync(h_table + at, pitch, d_table + at, pitch, width, r1 - r0,
                                 cudaMemcpyDeviceToHost, st));
    }
  }
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

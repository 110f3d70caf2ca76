// Path 2a without the big table: the scoring prologue of Video._index_changes
// (video.py:109-116, Bitmap.diff_weights screen.py:400-449) with every edit-distance entry
// evaluated from FACTOR tables that live in shared memory.
//
// An entry of the edit-distance table (make_data_tables.py:92-108) is the chain
//     F_t = min(F_{t+1} + S[a_t][b_t],  F_{t+2} + 1 if pixels (t, t+1) are a swapped pair)
// (csrc/iiv_tables.cu), i.e. the vector (F_t, F_{t+1}) is a 2x2 (min,+) matrix M_t times
// (F_{t+1}, F_{t+2}).  The product over a SEGMENT of pixels [p, q) depends on pixels p..q
// only, and those are fed by a 4-6 bit window of the masked value (colours.py:100-134:
// pixel t rotates dots t..t+3; screen.py:741-789 for which bits feed which HGR dots).  So
//     entry(x, y) = row0(T_0[x_0, y_0]) . T_1[x_1, y_1] . ... . T_{L-1}[..] . col(T_L[x_L, y_L])
// with T_k tabulated over pairs of k-th windows: 2^10..2^12 pairs each, 104 KiB per byte
// offset in both modes -- the two byte offsets of a bank fit the 227 KiB of shared
// memory of one SM.  A lookup is then 5 (DHGR) / 8 (HGR) shared-memory loads and a few
// packed 16-bit adds and mins instead of one 2-byte gather from a 512 MiB / 1 GiB table in
// HBM (93 B of DRAM traffic each, profiles/r02_gather_flavours.txt): the scorer stops being
// bound by DRAM row activations and the table is not needed at all.
//
// Results are bit-identical to the table path (tests/test_gpu_factored.py compares with
// iiv_score_frames and the oracle); tests/test_oracle_tables.py::test_factor_segments checks
// the windows against the oracle's pixel strings on the CPU.
#include "iiv_common.cuh"

namespace iiv {
namespace {

constexpr uint32_t kInf = 0x4000;   // "no swap": + any F (<= 18 * 255) stays below 0x8000

struct Lut {
  uint8_t s[256];
};

struct SegDesc {
  int p, q;        // pixels [p, q)
  uint32_t mask;   // the bits of the masked value that feed pixels p..min(q, n-1)
};

// Segments per (mode, window set).  HGR has one set per byte offset (the palette bit that
// shifts the body's dots is bit 10 at offset 0, bit 3 at offset 1); DHGR's dots are the
// value's bits whatever the offset.
template <int MODE, int WIN>
struct Chain;
template <>
struct Chain<IIV_MODE_HGR, 0> {
  static constexpr int kSegs = 8;
  __host__ __device__ static constexpr SegDesc seg(int k) {
    constexpr SegDesc s[kSegs] = {{0, 2, 0x041f},  {2, 4, 0x043a},   {4, 6, 0x0478},
                                  {6, 8, 0x04f0},  {8, 10, 0x05e0},  {10, 12, 0x07c0},
                                  {12, 15, 0x1f80}, {15, 18, 0x3f00}};
    return s[k];
  }
};
template <>
struct Chain<IIV_MODE_HGR, 1> {
  static constexpr int kSegs = 8;
  __host__ __device__ static constexpr SegDesc seg(int k) {
    constexpr SegDesc s[kSegs] = {{0, 2, 0x003f},  {2, 4, 0x007a},   {4, 6, 0x00f8},
                                  {6, 8, 0x01e8},  {8, 10, 0x03c8},  {10, 12, 0x0788},
                                  {12, 15, 0x1f08}, {15, 18, 0x3e08}};
    return s[k];
  }
};
template <>
struct Chain<IIV_MODE_DHGR, 0> {
  static constexpr int kSegs = 5;
  __host__ __device__ static constexpr SegDesc seg(int k) {
    constexpr SegDesc s[kSegs] = {
        {0, 2, 0x003f}, {2, 4, 0x00fc}, {4, 6, 0x03f0}, {6, 7, 0x07c0}, {7, 10, 0x1f80}};
    return s[k];
  }
};
template <int MODE>
__host__ __device__ constexpr int win_of(int o) {
  return MODE == IIV_MODE_HGR ? o : 0;
}

__host__ __device__ constexpr int popc_c(uint32_t x) {
  int n = 0;
  for (; x; x &= x - 1) ++n;
  return n;
}

// Table k of a chain: 4 bytes per window pair for the first (row 0 of T: T00 | T01 << 16) and
// the last (T . (0, INF): F_p | F_{p+1} << 16), 8 bytes for the ones between (T00 | T01 << 16,
// T10 | T11 << 16).
template <int MODE, int WIN>
__host__ __device__ constexpr uint32_t seg_entry_bytes(int k) {
  return (k == 0 || k == Chain<MODE, WIN>::kSegs - 1) ? 4u : 8u;
}
template <int MODE, int WIN>
__host__ __device__ constexpr uint32_t seg_offset(int k) {   // k == kSegs: the total
  uint32_t off = 0;
  for (int j = 0; j < k; ++j)
    off += seg_entry_bytes<MODE, WIN>(j) << (2 * popc_c(Chain<MODE, WIN>::seg(j).mask));
  return off;
}
template <int MODE>
constexpr uint32_t kOffsetBytes = seg_offset<MODE, 0>(Chain<MODE, 0>::kSegs);
static_assert(seg_offset<IIV_MODE_HGR, 1>(Chain<IIV_MODE_HGR, 1>::kSegs) == kOffsetBytes<IIV_MODE_HGR>,
              "both HGR window sets take the same room");
static_assert(2 * kOffsetBytes<IIV_MODE_HGR> <= 227 * 1024 &&
                  2 * kOffsetBytes<IIV_MODE_DHGR> <= 227 * 1024,
              "a bank's two byte offsets fit one SM's shared memory");

template <int MODE, int WIN, int K>
struct Seg {
  static constexpr SegDesc d = Chain<MODE, WIN>::seg(K);
  using E = Ext<d.mask>;
  static constexpr int bits = E::count;
  static constexpr uint32_t off = seg_offset<MODE, WIN>(K);
  static constexpr uint32_t entry = seg_entry_bytes<MODE, WIN>(K);
  __host__ __device__ static __forceinline__ uint32_t index(uint32_t x, uint32_t y) {
    return (E::ext(x) << bits) | E::ext(y);
  }
};

// ---- the factor tables -----------------------------------------------------------------
// One thread per (offset, segment, window pair): the segment's pixels straight from to_dots /
// nominal_pixel (no pixel-string table), then the recurrence from (F_q, F_{q+1}) = (0, INF)
// and (INF, 0): the two columns of T.
template <int MODE, int WIN, int K>
__device__ __forceinline__ void tabulate_segment(const uint8_t* S, int o, uint32_t idx,
                                                 unsigned char* tab) {
  using M = Mode<MODE>;
  using G = Seg<MODE, WIN, K>;
  constexpr int n = M::kDots, p = G::d.p, q = G::d.q;
  constexpr bool first = K == 0, last = K == Chain<MODE, WIN>::kSegs - 1;
  if (idx >= (1u << (2 * G::bits))) return;
  const uint32_t xi = idx >> G::bits, xj = idx & ((1u << G::bits) - 1u);
  const uint32_t da = to_dots<MODE>(G::E::dep(xi), o), db = to_dots<MODE>(G::E::dep(xj), o);
  uint32_t col[2][2];   // col[j] = (F_p, F_{p+1}) from unit vector j at (F_q, F_{q+1})
#pragma unroll
  for (int j = 0; j < (last ? 1 : 2); ++j) {
    uint32_t f1 = j == 0 ? 0u : kInf, f2 = j == 0 ? kInf : 0u;
#pragma unroll
    for (int t = q - 1; t >= p; --t) {
      const uint32_t a = nominal_pixel(da, t, M::phase(o)), b = nominal_pixel(db, t, M::phase(o));
      uint32_t cur = f1 + S[a * 16 + b];
      if (t + 1 < n) {
        const uint32_t na = nominal_pixel(da, t + 1, M::phase(o));
        const uint32_t nb = nominal_pixel(db, t + 1, M::phase(o));
        if (a == nb && na == b) cur = min(cur, f2 + 1u);
      }
      f2 = f1;
      f1 = min(cur, kInf);
    }
    col[j][0] = f1;
    col[j][1] = min(f2, kInf);
  }
  unsigned char* at = tab + G::off + (size_t)idx * G::entry;
  if (first)
    *reinterpret_cast<uint32_t*>(at) = col[0][0] | (col[1][0] << 16);   // row 0: T00, T01
  else if (last)
    *reinterpret_cast<uint32_t*>(at) = col[0][0] | (col[0][1] << 16);   // T . (0, INF)
  else
    *reinterpret_cast<uint2*>(at) =
        make_uint2(col[0][0] | (col[1][0] << 16), col[0][1] | (col[1][1] << 16));
}

template <int MODE, int WIN, int K = 0>
__device__ __forceinline__ void tabulate_dispatch(const uint8_t* S, int o, int k, uint32_t idx,
                                                  unsigned char* tab) {
  if constexpr (K < Chain<MODE, WIN>::kSegs) {
    if (k == K)
      tabulate_segment<MODE, WIN, K>(S, o, idx, tab);
    else
      tabulate_dispatch<MODE, WIN, K + 1>(S, o, k, idx, tab);
  }
}

// The blob: the tables of every byte offset, then a 16-byte trailer whose first word says
// whether the LUT's diagonal is zero -- then entry(x, x) = 0 and the scorer skips the lookup of
// a window that already shows what the target shows.
template <int MODE>
constexpr size_t kBlobBytes = (size_t)Mode<MODE>::kOffsets * kOffsetBytes<MODE> + 16;

template <int MODE>
__global__ void __launch_bounds__(256)
factor_prologue(const __grid_constant__ Lut lut, unsigned char* __restrict__ factors) {
  __shared__ uint8_t S[256];
  S[threadIdx.x] = lut.s[threadIdx.x];
  __syncthreads();
  const int o = blockIdx.z, k = blockIdx.y;
  const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
  if (idx == 0 && o == 0 && k == 0) {
    uint32_t zero = 1;
    for (int a = 0; a < 16; ++a) zero &= S[a * 17] == 0;
    *reinterpret_cast<uint4*>(factors + (size_t)Mode<MODE>::kOffsets * kOffsetBytes<MODE>) =
        make_uint4(zero, 0u, 0u, 0u);
  }
  unsigned char* tab = factors + (size_t)o * kOffsetBytes<MODE>;
  if (MODE == IIV_MODE_HGR && o == 1)
    tabulate_dispatch<MODE, win_of<MODE>(1)>(S, o, k, idx, tab);
  else
    tabulate_dispatch<MODE, 0>(S, o, k, idx, tab);
}

// ---- a lookup ----------------------------------------------------------------------------
// From the last segment backwards, v = F_p | F_{p+1} << 16.  A middle segment is two packed
// adds (row 0 + v, row 1 + v), two byte permutes that pair the halves up, one packed min.
template <int MODE, int WIN, int K>
__device__ __forceinline__ uint32_t chain_middle(const unsigned char* tab, uint32_t x, uint32_t y,
                                                 uint32_t v) {
  if constexpr (K >= 1) {
    using G = Seg<MODE, WIN, K>;
    const uint2 m = *reinterpret_cast<const uint2*>(tab + G::off + G::index(x, y) * 8u);
    const uint32_t t = m.x + v, u = m.y + v;
    v = __vminu2(__byte_perm(t, u, 0x5410), __byte_perm(t, u, 0x7632));
    return chain_middle<MODE, WIN, K - 1>(tab, x, y, v);
  } else {
    return v;
  }
}

template <int MODE, int WIN>
__device__ __forceinline__ uint32_t chain_lookup(const unsigned char* tab, uint32_t x, uint32_t y) {
  constexpr int L = Chain<MODE, WIN>::kSegs - 1;
  using GL = Seg<MODE, WIN, L>;
  using G0 = Seg<MODE, WIN, 0>;
  uint32_t v = *reinterpret_cast<const uint32_t*>(tab + GL::off + GL::index(x, y) * 4u);
  v = chain_middle<MODE, WIN, L - 1>(tab, x, y, v);
  const uint32_t t = *reinterpret_cast<const uint32_t*>(tab + G0::off + G0::index(x, y) * 4u) + v;
  return min(t & 0xffffu, t >> 16);
}

// ---- iiv_score_frames with the factors ---------------------------------------------------
// Persistent blocks, one per SM: a block keeps the factor tables of ONE bank's two byte
// offsets in shared memory and walks (frame, quarter-frame) items of that bank; a thread owns
// 4 adjacent packed columns of a page, as in score_frames_kernel: it builds their packed
// target words from the raw screen bytes of both banks, reads 32 bytes of packed source, does
// its 8 lookups and stores 32 B of diff weights and 32 B of priorities.
constexpr int kFactoredThreads = 1024;

template <int MODE>
__global__ void __launch_bounds__(kFactoredThreads, 1)
score_frames_factored_kernel(const uint64_t* __restrict__ src, size_t src_stride,
                             const uint8_t* __restrict__ tmain, const uint8_t* __restrict__ taux,
                             size_t mem_stride, const unsigned char* __restrict__ factors,
                             uint64_t* __restrict__ tpacked, int32_t* __restrict__ diff,
                             int32_t* __restrict__ prio, int zero_holes, int batch) {
  using M = Mode<MODE>;
  constexpr int kBanks = MODE == IIV_MODE_DHGR ? 2 : 1;
  constexpr uint32_t kTab = kOffsetBytes<MODE>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t tables_bar;
  const int bank = blockIdx.x % kBanks;   // index into diff / priority: 0 main, 1 aux
  // The bank's two tables come in as two TMA bulk copies (cp.async.bulk, 104 KiB each)
  // that complete on an mbarrier; the block's threads meanwhile fetch the screen bytes, source
  // words and priorities of their first item and wait only when they need a table entry.
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tables_bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                 "r"(2u * kTab)
                 : "memory");
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const unsigned char* g = factors + (size_t)byte_offset<MODE>(half, bank) * kTab;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + half * kTab);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(dst), "l"(g), "r"(kTab), "r"(bar)
          : "memory");
    }
  }
  const bool diag0 =
      *reinterpret_cast<const uint32_t*>(factors + (size_t)M::kOffsets * kTab) != 0u;
  __syncthreads();   // the barrier is initialised before anyone polls it
  bool tables_ready = false;

  const int group = threadIdx.x >> 8, t = threadIdx.x & 255;   // 4 groups of 256 threads
  const int n_items = 4 * batch, stride = (gridDim.x / kBanks) * 4;
  for (int item = (blockIdx.x / kBanks) * 4 + group; item < n_items; item += stride) {
    const size_t frame = item >> 2;
    const int c0 = (item & 3) * 1024 + 4 * t;   // first of 4 columns, one page
    const int col = c0 & 127;
    const uint8_t* mm = tmain + frame * mem_stride + 2 * c0;
    const uint8_t* am = MODE == IIV_MODE_DHGR ? taux + frame * mem_stride + 2 * c0 : nullptr;
    // bytes of columns c0-1 .. c0+4 (12 bytes per bank); outside the page: zeros, which give
    // the zero header / footer of screen.py:217, :224
    uint8_t mb[12], ab[12];
    {
      const uint2 w = *reinterpret_cast<const uint2*>(mm);
      const uchar2 p = col > 0 ? *reinterpret_cast<const uchar2*>(mm - 2) : make_uchar2(0, 0);
      const uchar2 n = col < 124 ? *reinterpret_cast<const uchar2*>(mm + 8) : make_uchar2(0, 0);
      mb[0] = p.x; mb[1] = p.y; mb[10] = n.x; mb[11] = n.y;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mb[2 + k] = (uint8_t)(w.x >> (8 * k));
        mb[6 + k] = (uint8_t)(w.y >> (8 * k));
      }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) ab[k] = 0;
    if (MODE == IIV_MODE_DHGR) {
      const uint2 w = *reinterpret_cast<const uint2*>(am);
      const uchar2 p = col > 0 ? *reinterpret_cast<const uchar2*>(am - 2) : make_uchar2(0, 0);
      const uchar2 n = col < 124 ? *reinterpret_cast<const uchar2*>(am + 8) : make_uchar2(0, 0);
      ab[0] = p.x; ab[1] = p.y; ab[10] = n.x; ab[11] = n.y;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ab[2 + k] = (uint8_t)(w.x >> (8 * k));
        ab[6 + k] = (uint8_t)(w.y >> (8 * k));
      }
    }
    uint64_t body[6];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      body[k] = body_of<MODE>(mb[2 * k], mb[2 * k + 1], ab[2 * k], ab[2 * k + 1]);
    uint64_t tp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      tp[k] = header_of<MODE>(body[k]) ^ body[k + 1] ^ footer_of<MODE>(body[k + 2]);
    const uint64_t* sp = src + frame * src_stride + c0;
    const ulonglong2 s01 = *reinterpret_cast<const ulonglong2*>(sp);
    const ulonglong2 s23 = *reinterpret_cast<const ulonglong2*>(sp + 2);
    const uint64_t sw[4] = {s01.x, s01.y, s23.x, s23.y};
    // the priorities stream in while the lookups run
    int4 pin[2];
    const size_t at = (frame * kBanks + bank) * 8192 + 2 * (size_t)c0;
    if (prio != nullptr) {
      const int4* pp = reinterpret_cast<const int4*>(prio + at);
      pin[0] = pp[0];
      pin[1] = pp[1];
    }
    if (tpacked != nullptr && bank == 0) {
      uint64_t* o = tpacked + frame * 4096 + c0;
      *reinterpret_cast<ulonglong2*>(o) = make_ulonglong2(tp[0], tp[1]);
      *reinterpret_cast<ulonglong2*>(o + 2) = make_ulonglong2(tp[2], tp[3]);
    }
    if (!tables_ready) {
      uint32_t done = 0;
      while (!done)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar)
            : "memory");
      tables_ready = true;
    }
    int32_t dw[8];
    const int o0 = byte_offset<MODE>(0, bank), o1 = byte_offset<MODE>(1, bank);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t x0 = mask_shift<MODE>(sw[k], o0), y0 = mask_shift<MODE>(tp[k], o0);
      const uint32_t x1 = mask_shift<MODE>(sw[k], o1), y1 = mask_shift<MODE>(tp[k], o1);
      dw[2 * k] = (diag0 && x0 == y0) ? 0 : (int32_t)chain_lookup<MODE, 0>(smem, x0, y0);
      dw[2 * k + 1] = (diag0 && x1 == y1) ? 0
                      : (int32_t)chain_lookup<MODE, MODE == IIV_MODE_HGR ? 1 : 0>(smem + kTab, x1, y1);
    }
    const bool hole = zero_holes && (col == 60 || col == 124);   // offsets 120..127, 248..255
    if (hole) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dw[k] = 0;                     // video.py:111
    }
    if (diff != nullptr) {
      int4* d = reinterpret_cast<int4*>(diff + at);
      d[0] = make_int4(dw[0], dw[1], dw[2], dw[3]);
      d[1] = make_int4(dw[4], dw[5], dw[6], dw[7]);
    }
    if (prio != nullptr) {
      int4* pp = reinterpret_cast<int4*>(prio + at);
      int32_t pv[8] = {pin[0].x, pin[0].y, pin[0].z, pin[0].w,
                       pin[1].x, pin[1].y, pin[1].z, pin[1].w};
#pragma unroll
      for (int k = 0; k < 8; ++k) pv[k] = (dw[k] == 0 ? 0 : pv[k]) + dw[k];   // video.py:115-116
      pp[0] = make_int4(pv[0], pv[1], pv[2], pv[3]);
      pp[1] = make_int4(pv[4], pv[5], pv[6], pv[7]);
    }
  }
}

inline bool mode_ok(int mode) { return mode == IIV_MODE_HGR || mode == IIV_MODE_DHGR; }

template <int MODE, int WIN>
void copy_segments(int* n_segments, int* p, int* q, uint32_t* masks) {
  *n_segments = Chain<MODE, WIN>::kSegs;
  for (int k = 0; k < Chain<MODE, WIN>::kSegs; ++k) {
    p[k] = Chain<MODE, WIN>::seg(k).p;
    q[k] = Chain<MODE, WIN>::seg(k).q;
    masks[k] = Chain<MODE, WIN>::seg(k).mask;
  }
}

template <int MODE>
int launch_prologue(const Lut& lut, unsigned char* d_factors, cudaStream_t st) {
  // the largest window pair count of any segment: 2^12
  factor_prologue<MODE><<<dim3((1u << 12) / 256, Chain<MODE, 0>::kSegs, Mode<MODE>::kOffsets),
                          256, 0, st>>>(lut, d_factors);
  IIV_LAUNCH_CHECK("factor_prologue");
  return 0;
}

template <int MODE>
int launch_score(const uint64_t* src, size_t src_stride, const uint8_t* tmain,
                 const uint8_t* taux, size_t mem_stride, const unsigned char* factors,
                 uint64_t* tpacked, int32_t* diff, int32_t* prio, int zero_holes, int batch,
                 cudaStream_t st) {
  constexpr int kBanks = MODE == IIV_MODE_DHGR ? 2 : 1;
  constexpr uint32_t smem = 2 * kOffsetBytes<MODE>;
  int dev = 0, sms = 0;
  IIV_CUDA(cudaGetDevice(&dev));
  IIV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  IIV_CUDA(cudaFuncSetAttribute(score_frames_factored_kernel<MODE>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sms - sms % kBanks;                      // one block per SM, banks in turn
  const int useful = kBanks * batch;                  // a block needs at least one quarter
  if (grid > useful * 4) grid = useful * 4;
  if (grid < kBanks) grid = kBanks;
  score_frames_factored_kernel<MODE><<<grid, kFactoredThreads, smem, st>>>(
      src, src_stride, tmain, taux, mem_stride, factors, tpacked, diff, prio, zero_holes, batch);
  IIV_LAUNCH_CHECK("score_frames_factored_kernel");
  return 0;
}

}  // namespace
}  // namespace iiv

using namespace iiv;

extern "C" size_t iiv_score_factors_bytes(int mode) {
  if (mode == IIV_MODE_HGR) return kBlobBytes<IIV_MODE_HGR>;
  if (mode == IIV_MODE_DHGR) return kBlobBytes<IIV_MODE_DHGR>;
  return 0;
}

extern "C" int iiv_score_factor_segments(int mode, int offset, int* n_segments, int* p, int* q,
                                         uint32_t* masks) {
  IIV_REQUIRE(n_segments && p && q && masks, "null pointer");
  if (mode == IIV_MODE_HGR && offset == 0)
    copy_segments<IIV_MODE_HGR, 0>(n_segments, p, q, masks);
  else if (mode == IIV_MODE_HGR && offset == 1)
    copy_segments<IIV_MODE_HGR, 1>(n_segments, p, q, masks);
  else if (mode == IIV_MODE_DHGR && offset >= 0 && offset < 4)
    copy_segments<IIV_MODE_DHGR, 0>(n_segments, p, q, masks);
  else
    IIV_REQUIRE(false, "bad mode %d / offset %d", mode, offset);
  return 0;
}

extern "C" int iiv_score_factors(int mode, const int32_t* h_lut, uint8_t* d_factors,
                                 void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(h_lut && d_factors && ((uintptr_t)d_factors % 16) == 0, "bad argument");
  Lut lut;
  for (int k = 0; k < 256; ++k) {
    if (h_lut[k] < 0 || h_lut[k] > 255) {
      set_error("substitution cost %d at [%d][%d] outside 0..255", h_lut[k], k >> 4, k & 15);
      return IIV_E_OVERFLOW;
    }
    lut.s[k] = (uint8_t)h_lut[k];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == IIV_MODE_HGR) return launch_prologue<IIV_MODE_HGR>(lut, d_factors, st);
  return launch_prologue<IIV_MODE_DHGR>(lut, d_factors, st);
}

extern "C" int iiv_score_frames_factored(int mode, const uint64_t* d_source_packed,
                                         size_t source_stride, const uint8_t* d_target_main,
                                         const uint8_t* d_target_aux, size_t mem_stride,
                                         const uint8_t* d_factors, uint64_t* d_target_packed,
                                         int32_t* d_diff, int32_t* d_priority, int zero_holes,
                                         int batch, void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(d_source_packed && d_target_main && d_factors && batch >= 0 && batch <= (1 << 20),
              "bad argument");
  IIV_REQUIRE(mode == IIV_MODE_HGR || d_target_aux, "DHGR needs aux memory");
  IIV_REQUIRE(mem_stride % 8 == 0 && source_stride % 2 == 0,
              "strides must keep 8/16-byte alignment");
  IIV_REQUIRE(((uintptr_t)d_target_main % 8) == 0 && ((uintptr_t)d_target_aux % 8) == 0 &&
                  ((uintptr_t)d_source_packed % 16) == 0 && ((uintptr_t)d_target_packed % 16) == 0 &&
                  ((uintptr_t)d_diff % 16) == 0 && ((uintptr_t)d_priority % 16) == 0 &&
                  ((uintptr_t)d_factors % 16) == 0,
              "misaligned buffer");
  if (batch == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == IIV_MODE_HGR)
    return launch_score<IIV_MODE_HGR>(d_source_packed, source_stride, d_target_main, d_target_aux,
                                      mem_stride, d_factors, d_target_packed, d_diff, d_priority,
                                      zero_holes, batch, st);
  return launch_score<IIV_MODE_DHGR>(d_source_packed, source_stride, d_target_main, d_target_aux,
                                     mem_stride, d_factors, d_target_packed, d_diff, d_priority,
                                     zero_holes, batch, st);
}

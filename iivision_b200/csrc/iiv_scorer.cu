// Path 2a: the data-parallel scoring primitives of transcoder/screen.py.
//
// All of them are byte/integer work bounded by memory traffic: packed screens
// are 32 KiB, the edit-distance tables 512 MiB (DHGR) / 1 GiB (HGR) and are hit
// by sector-granular random 2-byte gathers.  Threads are laid out so that the
// packed words and the int32 outputs move coalesced, and every thread keeps
// several independent table gathers in flight.
#include "iiv_common.cuh"

namespace iiv {
namespace {

// ---- Bitmap._pack (screen.py:207-226) -----------------------------------------
// One block per screen: 4096 columns, 256 threads x 16 columns.  Bodies are
// staged in shared memory so header/footer come from the neighbours without
// re-reading global memory.
template <int MODE>
__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t* __restrict__ main_mem,
            const uint8_t* __restrict__ aux_mem, size_t mem_stride,
            uint64_t* __restrict__ packed) {
  __shared__ uint64_t body[32 * 128];
  const uint8_t* mm = main_mem + (size_t)blockIdx.x * mem_stride;
  const uint8_t* am =
      MODE == IIV_MODE_DHGR ? aux_mem + (size_t)blockIdx.x * mem_stride : nullptr;
  uint64_t* out = packed + (size_t)blockIdx.x * 4096;
  for (int c = threadIdx.x; c < 4096; c += 256) {
    const uchar2 m = reinterpret_cast<const uchar2*>(mm)[c];
    uchar2 a = make_uchar2(0, 0);
    if (MODE == IIV_MODE_DHGR) a = reinterpret_cast<const uchar2*>(am)[c];
    body[c] = body_of<MODE>(m.x, m.y, a.x, a.y);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 4096; c += 256) {
    const int col = c & 127;
    // header[:,0] = 0 and footer[:,-1] = 0 (screen.py:217, :224); interior
    // columns do leak across the 40-byte row segments, as in the reference.
    const uint64_t h = col > 0 ? header_of<MODE>(body[c - 1]) : 0;
    const uint64_t f = col < 127 ? footer_of<MODE>(body[c + 1]) : 0;
    out[c] = h ^ body[c] ^ f;
  }
}

template <int MODE>
__global__ void mask_shift_kernel(int o, const uint64_t* __restrict__ in,
                                  uint64_t* __restrict__ out, size_t n) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = mask_shift<MODE>(in[k], o);
}

template <int MODE>
__global__ void masked_update_kernel(int o, const uint64_t* __restrict__ in,
                                     uint32_t v, uint64_t* __restrict__ out,
                                     size_t n) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = masked_update<MODE>(o, in[k], v);
}

template <int MODE>
__global__ void column_part_kernel(int part, const uint64_t* __restrict__ in,
                                   uint64_t* __restrict__ out, size_t n) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint64_t w = in[k];
  uint64_t r;
  if (part == IIV_PART_HEADER)
    r = header_of<MODE>(w);
  else if (part == IIV_PART_FOOTER)
    r = footer_of<MODE>(w);
  else if (part == IIV_PART_BODY)
    r = w & keep_low_mask<MODE>() & keep_high_mask<MODE>();
  else
    r = double_pixels((uint32_t)w & 0x7fu);
  out[k] = r;
}

template <int MODE>
__global__ void fix_column_kernel(int side, const uint64_t* __restrict__ neighbour,
                                  const uint64_t* __restrict__ column,
                                  uint64_t* __restrict__ out, size_t n) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  out[k] = side == 0 ? (neighbour[k] & keep_low_mask<MODE>()) ^ footer_of<MODE>(column[k])
                     : (neighbour[k] & keep_high_mask<MODE>()) ^ header_of<MODE>(column[k]);
}

// Bitmap._fix_array_neighbours (screen.py:322-341): np.roll wraps inside a row.
template <int MODE>
__global__ void __launch_bounds__(128)
fix_neighbours_kernel(int o, uint64_t* __restrict__ rows) {
  __shared__ uint64_t r[128];
  uint64_t* row = rows + (size_t)blockIdx.x * 128;
  const int c = threadIdx.x;
  r[c] = row[c];
  __syncthreads();
  if (o == 0)
    row[c] = (r[c] & keep_low_mask<MODE>()) ^ footer_of<MODE>(r[(c + 1) & 127]);
  else if (o == Mode<MODE>::kOffsets - 1)
    row[c] = (r[c] & keep_high_mask<MODE>()) ^ header_of<MODE>(r[(c + 127) & 127]);
}

// ---- Bitmap._diff_weights / _diff_weights_page (screen.py:409-494) ------------
// One thread per packed column: it owns both interleaved outputs (even/odd page
// offsets = the two byte offsets of the bank), so one 64-bit load of source and
// target feeds two gathers and one 8-byte store.  The neighbour fix-up of the
// content variant never touches the masked window being scored (it rewrites
// footer bits for o == 0 and header bits for o == last, both outside that
// offset's mask), so it is elided here; iiv_fix_array_neighbours exists for
// callers who want the intermediate array.
template <int MODE>
__global__ void __launch_bounds__(256)
diff_weights_kernel(int is_aux, const uint64_t* __restrict__ src,
                    const uint64_t* __restrict__ tgt, int content,
                    const uint16_t* __restrict__ table,
                    int32_t* __restrict__ out, size_t n_cols) {
  using M = Mode<MODE>;
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  const uint64_t s = src[c], t = tgt[c];
  int2 res;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int o = byte_offset<MODE>(half, is_aux);
    const uint64_t cmp = content >= 0 ? masked_update<MODE>(o, s, (uint32_t)content) : s;
    const uint32_t x = mask_shift<MODE>(cmp, o);
    const uint32_t y = mask_shift<MODE>(t, o);
    const uint16_t d = (uint16_t)ldg_table(table + ((size_t)o << (2 * M::kBits)) +
                                           ((size_t)x << M::kBits) + y);
    (half ? res.y : res.x) = d;
  }
  reinterpret_cast<int2*>(out)[c] = res;
}

// ---- Video._index_changes' scoring prologue for a batch of frames ---------------------
// video.py:109-116 (diff_weights of the target against the source bitmap, holes zeroed,
// priorities folded) fused with Bitmap._pack of the target (screen.py:207-226), both banks
// of a DHGR frame in one pass.  A block owns 8 pages (1024 packed columns) of one frame, a
// thread 4 adjacent columns: it builds their packed target words from the raw screen bytes
// (its own 8 bytes of each bank plus the neighbouring column on either side for header and
// footer), reads 32 bytes of packed source, and issues all of its table gathers -- 4
// columns x 2 byte offsets x banks = 16 (DHGR) / 8 (HGR) independent loads -- before it
// touches any result.  Outputs leave as 16-byte stores: 32 B of packed target, 32 B of
// diff weights and 32 B of priorities per thread and bank.
template <int MODE>
__global__ void __launch_bounds__(256)
score_frames_kernel(const uint64_t* __restrict__ src, size_t src_stride,
                    const uint8_t* __restrict__ tmain, const uint8_t* __restrict__ taux,
                    size_t mem_stride, const uint16_t* __restrict__ table,
                    uint64_t* __restrict__ tpacked, int32_t* __restrict__ diff,
                    int32_t* __restrict__ prio, int zero_holes) {
  using M = Mode<MODE>;
  constexpr int kBanks = MODE == IIV_MODE_DHGR ? 2 : 1;
  const size_t frame = blockIdx.y;
  const int c0 = blockIdx.x * 1024 + 4 * threadIdx.x;   // first of 4 columns, one page
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
  for (int k = 0; k < 4; ++k) {
    // header[:,0] = 0, footer[:,-1] = 0: the zero bodies above stand in at the page edges,
    // but a zero BODY is not a zero header for HGR only if bits leak -- header_of(0) == 0
    tp[k] = header_of<MODE>(body[k]) ^ body[k + 1] ^ footer_of<MODE>(body[k + 2]);
  }
  const uint64_t* sp = src + frame * src_stride + c0;
  const ulonglong2 s01 = *reinterpret_cast<const ulonglong2*>(sp);
  const ulonglong2 s23 = *reinterpret_cast<const ulonglong2*>(sp + 2);
  const uint64_t sw[4] = {s01.x, s01.y, s23.x, s23.y};
  if (tpacked != nullptr) {
    uint64_t* o = tpacked + frame * 4096 + c0;
    *reinterpret_cast<ulonglong2*>(o) = make_ulonglong2(tp[0], tp[1]);
    *reinterpret_cast<ulonglong2*>(o + 2) = make_ulonglong2(tp[2], tp[3]);
  }
  uint32_t idx[kBanks][8];
#pragma unroll
  for (int b = 0; b < kBanks; ++b)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int o = byte_offset<MODE>(half, b);
        idx[b][2 * k + half] = ((uint32_t)o << (2 * M::kBits)) +
                               (mask_shift<MODE>(sw[k], o) << M::kBits) +
                               mask_shift<MODE>(tp[k], o);
      }
  int32_t dw[kBanks][8];
#pragma unroll
  for (int b = 0; b < kBanks; ++b)
#pragma unroll
    for (int k = 0; k < 8; ++k) dw[b][k] = (int32_t)ldg_table(table + idx[b][k]);
  // the priorities stream in while the gathers are in flight
  int4 pin[kBanks][2];
  if (prio != nullptr) {
#pragma unroll
    for (int b = 0; b < kBanks; ++b) {
      const int4* pp =
          reinterpret_cast<const int4*>(prio + (frame * kBanks + b) * 8192 + 2 * (size_t)c0);
      pin[b][0] = pp[0];
      pin[b][1] = pp[1];
    }
  }
  const bool hole = zero_holes && (col == 60 || col == 124);   // offsets 120..127, 248..255
#pragma unroll
  for (int b = 0; b < kBanks; ++b) {
    if (hole) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dw[b][k] = 0;                // video.py:111
    }
    const size_t at = (frame * kBanks + b) * 8192 + 2 * (size_t)c0;
    if (diff != nullptr) {
      int4* d = reinterpret_cast<int4*>(diff + at);
      d[0] = make_int4(dw[b][0], dw[b][1], dw[b][2], dw[b][3]);
      d[1] = make_int4(dw[b][4], dw[b][5], dw[b][6], dw[b][7]);
    }
    if (prio != nullptr) {
      int4* pp = reinterpret_cast<int4*>(prio + at);
      const int4 p0 = pin[b][0], p1 = pin[b][1];
      int32_t pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
      for (int k = 0; k < 8; ++k)
        pv[k] = (dw[b][k] == 0 ? 0 : pv[k]) + dw[b][k];        // video.py:115-116
      pp[0] = make_int4(pv[0], pv[1], pv[2], pv[3]);
      pp[1] = make_int4(pv[4], pv[5], pv[6], pv[7]);
    }
  }
}

// compute_delta_page (screen.py:525-547) for one (page, content).
template <int MODE>
__global__ void __launch_bounds__(128)
delta_page_kernel(int is_aux, const uint64_t* __restrict__ tgt_row, int content,
                  const int32_t* __restrict__ diff_row,
                  const uint16_t* __restrict__ table, int32_t* __restrict__ out) {
  using M = Mode<MODE>;
  const int c = threadIdx.x;
  const uint64_t t = tgt_row[c];
  const int2 dr = reinterpret_cast<const int2*>(diff_row)[c];
  int2 res;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int o = byte_offset<MODE>(half, is_aux);
    const uint32_t x = mask_shift<MODE>(masked_update<MODE>(o, t, (uint32_t)content), o);
    const uint32_t y = mask_shift<MODE>(t, o);
    const int32_t d = (int32_t)ldg_table(table + ((size_t)o << (2 * M::kBits)) +
                                         ((size_t)x << M::kBits) + y);
    (half ? res.y : res.x) = d - (half ? dr.y : dr.x);
  }
  reinterpret_cast<int2*>(out)[c] = res;
}

// All (page, content) new-diff rows of a target bank.  grid = (32 pages, batch),
// block = 128 columns; each thread walks the contents for its column so that the
// 2 * n_content gathers of one column (which share header/footer bits and hence
// table neighbourhood) are issued back to back.
template <int MODE>
__global__ void __launch_bounds__(128)
delta_rows_kernel(int is_aux, const uint64_t* __restrict__ tgt,
                  const uint16_t* __restrict__ table, uint16_t* __restrict__ out) {
  using M = Mode<MODE>;
  const int page = blockIdx.x, c = threadIdx.x;
  const size_t screen = blockIdx.y;
  const uint64_t t = tgt[screen * 4096 + page * 128 + c];
  uint16_t* dst = out + ((screen * 32 + page) * M::kContents) * 256 + 2 * c;
  const int o0 = byte_offset<MODE>(0, is_aux), o1 = byte_offset<MODE>(1, is_aux);
  const uint16_t* t0 = table + ((size_t)o0 << (2 * M::kBits)) + mask_shift<MODE>(t, o0);
  const uint16_t* t1 = table + ((size_t)o1 << (2 * M::kBits)) + mask_shift<MODE>(t, o1);
#pragma unroll 8
  for (int content = 0; content < M::kContents; ++content) {
    const uint32_t x0 = mask_shift<MODE>(masked_update<MODE>(o0, t, content), o0);
    const uint32_t x1 = mask_shift<MODE>(masked_update<MODE>(o1, t, content), o1);
    // symmetric table: T[x][y] == T[y][x]; index as (x << bits) + y like the
    // reference (source-with-content in the high half).
    const uint32_t d0 = ldg_table(t0 + ((size_t)x0 << M::kBits));
    const uint32_t d1 = ldg_table(t1 + ((size_t)x1 << M::kBits));
    *reinterpret_cast<uint32_t*>(dst + (size_t)content * 256) = d0 | (d1 << 16);
  }
}

template <int MODE>
__global__ void pair_difference_kernel(int o, const uint64_t* __restrict__ old_packed,
                                       const uint8_t* __restrict__ content,
                                       const uint16_t* __restrict__ table,
                                       uint16_t* __restrict__ out, size_t n) {
  using M = Mode<MODE>;
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint64_t w = old_packed[k];
  const uint32_t oldp = mask_shift<MODE>(w, o);
  const uint32_t newp = mask_shift<MODE>(masked_update<MODE>(o, w, content[k]), o);
  out[k] = (uint16_t)ldg_table(table + ((size_t)o << (2 * M::kBits)) + ((size_t)oldp << M::kBits) + newp);
}

struct Stores {
  int32_t v[4 * 64];
};

// Bitmap.apply (screen.py:256-293) for up to 64 stores, in order, one thread.
template <int MODE>
__global__ void apply_kernel(uint64_t* packed, uint8_t* main_mem, uint8_t* aux_mem,
                             Stores st, int n) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int k = 0; k < n; ++k) {
    const int page = st.v[4 * k], offset = st.v[4 * k + 1];
    const int is_aux = st.v[4 * k + 2];
    const uint32_t value = (uint32_t)st.v[4 * k + 3] & 0xffu;
    const int o = byte_offset<MODE>(offset, is_aux);
    const int c = offset >> 1;
    uint64_t* row = packed + page * 128;
    row[c] = masked_update<MODE>(o, row[c], value);
    if (o == 0 && c > 0)
      row[c - 1] = (row[c - 1] & keep_low_mask<MODE>()) ^ footer_of<MODE>(row[c]);
    else if (o == Mode<MODE>::kOffsets - 1 && c < 127)
      row[c + 1] = (row[c + 1] & keep_high_mask<MODE>()) ^ header_of<MODE>(row[c]);
    (is_aux ? aux_mem : main_mem)[page * 256 + offset] = (uint8_t)value;
  }
}

inline bool mode_ok(int mode) { return mode == IIV_MODE_HGR || mode == IIV_MODE_DHGR; }
inline int n_offsets(int mode) { return mode == IIV_MODE_HGR ? 2 : 4; }

#define IIV_DISPATCH(mode, KERNEL, grid, block, st, ...)                    \
  do {                                                                      \
    if ((mode) == IIV_MODE_HGR)                                             \
      KERNEL<IIV_MODE_HGR><<<grid, block, 0, st>>>(__VA_ARGS__);            \
    else                                                                    \
      KERNEL<IIV_MODE_DHGR><<<grid, block, 0, st>>>(__VA_ARGS__);           \
    IIV_LAUNCH_CHECK(#KERNEL);                                              \
  } while (0)

}  // namespace
}  // namespace iiv

using namespace iiv;

extern "C" int iiv_pack(int mode, const uint8_t* d_main, const uint8_t* d_aux,
                        size_t mem_stride, uint64_t* d_packed, int batch,
                        void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(d_main && d_packed && batch >= 0, "bad argument");
  IIV_REQUIRE(mode == IIV_MODE_HGR || d_aux, "DHGR needs aux memory");
  IIV_REQUIRE(mem_stride % 2 == 0, "mem_stride must be even");
  if (batch == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  IIV_DISPATCH(mode, pack_kernel, batch, 256, st, d_main, d_aux, mem_stride, d_packed);
  return 0;
}

extern "C" int iiv_mask_and_shift(int mode, int byte_offset, const uint64_t* d_in,
                                  uint64_t* d_out, size_t n, void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(byte_offset >= 0 && byte_offset < n_offsets(mode), "bad byte_offset %d", byte_offset);
  IIV_REQUIRE(d_in && d_out, "null pointer");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  IIV_DISPATCH(mode, mask_shift_kernel, (unsigned)((n + 255) / 256), 256, st,
               byte_offset, d_in, d_out, n);
  return 0;
}

extern "C" int iiv_masked_update(int mode, int byte_offset, const uint64_t* d_old,
                                 uint8_t value, uint64_t* d_new, size_t n,
                                 void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(byte_offset >= 0 && byte_offset < n_offsets(mode), "bad byte_offset %d", byte_offset);
  IIV_REQUIRE(d_old && d_new, "null pointer");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  IIV_DISPATCH(mode, masked_update_kernel, (unsigned)((n + 255) / 256), 256, st,
               byte_offset, d_old, (uint32_t)value, d_new, n);
  return 0;
}

extern "C" int iiv_column_part(int mode, int part, const uint64_t* d_in, uint64_t* d_out,
                               size_t n, void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(part >= IIV_PART_HEADER && part <= IIV_PART_DOUBLE, "bad part %d", part);
  IIV_REQUIRE(!(part == IIV_PART_DOUBLE && mode != IIV_MODE_HGR), "_double_pixels is HGR only");
  IIV_REQUIRE(d_in && d_out, "null pointer");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  IIV_DISPATCH(mode, column_part_kernel, (unsigned)((n + 255) / 256), 256, st, part, d_in,
               d_out, n);
  return 0;
}

extern "C" int iiv_fix_column(int mode, int side, const uint64_t* d_neighbour,
                              const uint64_t* d_column, uint64_t* d_out, size_t n,
                              void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(side == 0 || side == 1, "bad side %d", side);
  IIV_REQUIRE(d_neighbour && d_column && d_out, "null pointer");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  IIV_DISPATCH(mode, fix_column_kernel, (unsigned)((n + 255) / 256), 256, st, side,
               d_neighbour, d_column, d_out, n);
  return 0;
}

extern "C" int iiv_fix_array_neighbours(int mode, int byte_offset, uint64_t* d_rows,
                                        int n_rows, void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(byte_offset >= 0 && byte_offset < n_offsets(mode), "bad byte_offset %d", byte_offset);
  IIV_REQUIRE(d_rows && n_rows >= 0, "bad argument");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  IIV_DISPATCH(mode, fix_neighbours_kernel, n_rows, 128, st, byte_offset, d_rows);
  return 0;
}

extern "C" int iiv_diff_weights(int mode, int is_aux, const uint64_t* d_source_packed,
                                const uint64_t* d_target_packed, int content,
                                const uint16_t* d_table, int32_t* d_out, int batch,
                                void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(!(mode == IIV_MODE_HGR && is_aux), "HGR has no aux bank");
  IIV_REQUIRE(d_source_packed && d_target_packed && d_table && d_out && batch >= 0, "bad argument");
  IIV_REQUIRE(content < 256, "bad content %d", content);
  if (batch == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_cols = (size_t)batch * 4096;
  IIV_DISPATCH(mode, diff_weights_kernel, (unsigned)((n_cols + 255) / 256), 256, st,
               is_aux, d_source_packed, d_target_packed, content, d_table, d_out, n_cols);
  return 0;
}

extern "C" int iiv_score_frames(int mode, const uint64_t* d_source_packed, size_t source_stride,
                                const uint8_t* d_target_main, const uint8_t* d_target_aux,
                                size_t mem_stride, const uint16_t* d_table,
                                uint64_t* d_target_packed, int32_t* d_diff, int32_t* d_priority,
                                int zero_holes, int batch, void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(d_source_packed && d_target_main && d_table && batch >= 0 && batch <= 65535,
              "bad argument");
  IIV_REQUIRE(mode == IIV_MODE_HGR || d_target_aux, "DHGR needs aux memory");
  IIV_REQUIRE(mem_stride % 8 == 0 && source_stride % 2 == 0, "strides must keep 8/16-byte alignment");
  IIV_REQUIRE(((uintptr_t)d_target_main % 8) == 0 && ((uintptr_t)d_target_aux % 8) == 0 &&
                  ((uintptr_t)d_source_packed % 16) == 0 && ((uintptr_t)d_target_packed % 16) == 0 &&
                  ((uintptr_t)d_diff % 16) == 0 && ((uintptr_t)d_priority % 16) == 0,
              "misaligned buffer");
  if (batch == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  IIV_DISPATCH(mode, score_frames_kernel, dim3(4, batch), 256, st, d_source_packed,
               source_stride, d_target_main, d_target_aux, mem_stride, d_table,
               d_target_packed, d_diff, d_priority, zero_holes);
  return 0;
}

extern "C" int iiv_diff_weights_page(int mode, int is_aux, const uint64_t* d_source_rows,
                                     const uint64_t* d_target_rows, int content,
                                     const uint16_t* d_table, int32_t* d_out,
                                     int n_rows, void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(!(mode == IIV_MODE_HGR && is_aux), "HGR has no aux bank");
  IIV_REQUIRE(d_source_rows && d_target_rows && d_table && d_out && n_rows >= 0, "bad argument");
  IIV_REQUIRE(content < 256, "bad content %d", content);
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_cols = (size_t)n_rows * 128;
  IIV_DISPATCH(mode, diff_weights_kernel, (unsigned)((n_cols + 255) / 256), 256, st,
               is_aux, d_source_rows, d_target_rows, content, d_table, d_out, n_cols);
  return 0;
}

extern "C" int iiv_compute_delta_page(int mode, int is_aux, const uint64_t* d_target_packed,
                                      int page, int content, const int32_t* d_diff_row,
                                      const uint16_t* d_table, int32_t* d_out,
                                      void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(!(mode == IIV_MODE_HGR && is_aux), "HGR has no aux bank");
  IIV_REQUIRE(d_target_packed && d_diff_row && d_table && d_out, "null pointer");
  IIV_REQUIRE(page >= 0 && page < 32 && content >= 0 && content < 256, "bad page/content");
  cudaStream_t st = (cudaStream_t)stream;
  IIV_DISPATCH(mode, delta_page_kernel, 1, 128, st, is_aux,
               d_target_packed + (size_t)page * 128, content, d_diff_row, d_table, d_out);
  return 0;
}

extern "C" int iiv_delta_rows(int mode, int is_aux, const uint64_t* d_target_packed,
                              const uint16_t* d_table, uint16_t* d_out, int batch,
                              void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(!(mode == IIV_MODE_HGR && is_aux), "HGR has no aux bank");
  IIV_REQUIRE(d_target_packed && d_table && d_out && batch >= 0 && batch <= 65535, "bad argument");
  if (batch == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  IIV_DISPATCH(mode, delta_rows_kernel, dim3(32, batch), 128, st, is_aux,
               d_target_packed, d_table, d_out);
  return 0;
}

extern "C" int iiv_byte_pair_difference(int mode, int byte_offset,
                                        const uint64_t* d_old_packed,
                                        const uint8_t* d_content,
                                        const uint16_t* d_table, uint16_t* d_out,
                                        size_t n, void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(byte_offset >= 0 && byte_offset < n_offsets(mode), "bad byte_offset %d", byte_offset);
  IIV_REQUIRE(d_old_packed && d_content && d_table && d_out, "null pointer");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  IIV_DISPATCH(mode, pair_difference_kernel, (unsigned)((n + 255) / 256), 256, st,
               byte_offset, d_old_packed, d_content, d_table, d_out, n);
  return 0;
}

extern "C" int iiv_apply(int mode, uint64_t* d_packed, uint8_t* d_main, uint8_t* d_aux,
                         const int32_t* h_stores, int n, void* stream) {
  IIV_REQUIRE(mode_ok(mode), "bad mode %d", mode);
  IIV_REQUIRE(d_packed && d_main && h_stores && n >= 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < n; base += 64) {
    Stores s;
    const int m = n - base < 64 ? n - base : 64;
    for (int k = 0; k < m; ++k) {
      const int32_t* q = h_stores + 4 * (base + k);
      IIV_REQUIRE(q[0] >= 0 && q[0] < 32 && q[1] >= 0 && q[1] < 256, "store %d out of range", base + k);
      IIV_REQUIRE(!(q[2] && (mode == IIV_MODE_HGR || !d_aux)), "aux store without aux bank");
      for (int e = 0; e < 4; ++e) s.v[4 * k + e] = q[e];
    }
    IIV_DISPATCH(mode, apply_kernel, 1, 32, st, d_packed, d_main, d_aux, s, m);
  }
  return 0;
}

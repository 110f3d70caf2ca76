// "Next" row N1: the compressed .npz the reference writes for every table
// (make_data_tables.py:186-188, np.savez_compressed) -- the deflate runs on the device, so
// that only compressed bytes cross PCIe and the host writes ZIP records.
//
// A table is cut into blocks of 32 KiB (16 384 entries: one HGR row, two DHGR rows).  Each
// becomes one dynamic-Huffman deflate block (RFC 1951) followed by an empty stored block
// that byte-aligns it, so blocks are independent and any inflater reads their
// concatenation.  Each of a block's 128 threads parses 256 bytes greedily, byte by byte:
//   * the longest match of >= 3 bytes among a few candidate distances -- whole entries
//     back, so only same-half bytes are compared: the previous entry (runs: the zeros
//     above the diagonal) and fixed column distances where these tables repeat (Cands<>:
//     columns that differ in one bit which the row's string happens not to care about --
//     or not at all: HGR maps 16 384 masked values onto 10 710 pixel strings, and the
//     columns with equal strings sit 4 or 2048 apart).  Matches may start and end on
//     either byte of an entry: a lone repeated entry between two equal high bytes is the
//     most frequent match of all;
//   * else one literal byte.
// The Huffman codes come from the host (deflate.py), built from the histogram the survey
// kernel takes over a sample of the blocks; a block that would not shrink is stored.
// The survey kernel also takes the CRC-32 of every block for the ZIP member's checksum.
#include "iiv_common.cuh"

namespace iiv {
namespace {

constexpr int kBlockEntries = 16384;             // 32 KiB of table per deflate block
constexpr int kBlockBytes = 2 * kBlockEntries;
constexpr int kDThreads = 128;
constexpr int kPerThread = kBlockEntries / kDThreads;   // 128 entries = 256 bytes
constexpr int kStride = kBlockBytes + 1024;      // room per block in the scratch buffer
constexpr int kOutWords = kStride / 4;
constexpr int kHist = 320;                       // 286 literal/length + 30 distance (+ pad)
constexpr int kCodeWords = 317 + 96;             // deflate.CodeTable.words()
// entry p lives at p + 2 * (p / 128): a thread's 128 entries start one bank further along
// than its neighbour's, so the per-thread sequential walks do not collide
__device__ __forceinline__ int padded(int p) { return p + ((p >> 7) << 1); }
constexpr int kPaddedEntries = kBlockEntries + 2 * kDThreads;

// RFC 1951 3.2.5: match length 3..258 -> symbol 257..285, extra bits
__device__ __forceinline__ void length_code(int len, int& sym, int& extra, int& nbits) {
  if (len == 258) {
    sym = 285; extra = 0; nbits = 0;
    return;
  }
  const int l = len - 3;                 // 0..254
  if (l < 8) {
    sym = 257 + l; extra = 0; nbits = 0;
    return;
  }
  const int hb = 31 - __clz(l);          // 3..7
  nbits = hb - 2;
  sym = 257 + 4 * nbits + 4 + ((l >> nbits) & 3);
  extra = l & ((1 << nbits) - 1);
}

// distance 1..32768 -> symbol 0..29, extra bits
__device__ __forceinline__ void distance_code(int dist, int& sym, int& extra, int& nbits) {
  const int d = dist - 1;                // 0..32767
  if (d < 4) {
    sym = d; extra = 0; nbits = 0;
    return;
  }
  const int hb = 31 - __clz(d);          // 2..14
  nbits = hb - 1;
  sym = 2 * hb + ((d >> nbits) & 1);
  extra = d & ((1 << nbits) - 1);
}

struct BlockGeom {
  int o;            // byte offset (table slice) of the block
  uint32_t col0;    // column of the block's first entry
  uint32_t n_mask;  // columns per row - 1
};

template <int MODE>
__device__ __forceinline__ BlockGeom geom(uint32_t block) {
  using M = Mode<MODE>;
  const uint64_t first = (uint64_t)block * kBlockEntries;
  BlockGeom g;
  g.o = (int)(first >> (2 * M::kBits));
  g.n_mask = (1u << M::kBits) - 1u;
  g.col0 = (uint32_t)first & g.n_mask;
  return g;
}

// Stage one block in shared memory (padded layout), coalesced 16-byte loads.
__device__ __forceinline__ void load_block(const uint16_t* __restrict__ table, uint32_t block,
                                           uint16_t* e) {
  const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)block * kBlockEntries);
  for (int k = threadIdx.x; k < kBlockBytes / 16; k += kDThreads) {
    const uint4 v = __ldg(src + k);
    // 8 entries, never straddling a 128-entry group
    uint32_t* dst = reinterpret_cast<uint32_t*>(e + padded(8 * k));
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
}

// Fixed candidate distances, in entries, ascending: where zlib -6 finds its long and its
// frequent matches in these tables (distance histogram of its token stream over rows spread
// through the table: 1 = runs, above all the zeros of the triangular layout; 4 and
// 48/96/192 = columns differing in one low / middle bit; 2048 = bit 11, long runs).  DHGR
// has no such favourite beyond the runs.
template <int MODE> struct Cands;
template <> struct Cands<IIV_MODE_HGR> {
  static constexpr int kN = 6;
  __device__ static constexpr int at(int k) {
    return k == 0 ? 1 : k == 1 ? 4 : k == 2 ? 48 : k == 3 ? 96 : k == 4 ? 192 : 2048;
  }
};
template <> struct Cands<IIV_MODE_DHGR> {
  static constexpr int kN = 5;
  __device__ static constexpr int at(int k) {
    return k == 0 ? 1 : k == 1 ? 4 : k == 2 ? 128 : k == 3 ? 512 : 2048;
  }
};

constexpr uint32_t kNoMatch = 0x0101u;   // an XOR with both bytes different

// XOR of entry q with the entry d back, or kNoMatch when there is none in this block.
__device__ __forceinline__ uint32_t diff_at(const uint16_t* e, int q, int d) {
  return d <= q ? (uint32_t)e[padded(q)] ^ (uint32_t)e[padded(q - d)] : kNoMatch;
}

// Bytes that repeat the bytes 2 * d earlier, from half `s` (0 = low byte) of entry p up to
// the end of the thread's chunk.  Distances are whole entries, so a byte is always compared
// with the same half of an earlier entry: one XOR of two entries answers for both bytes.
__device__ __forceinline__ int match_bytes(const uint16_t* e, int p, int s, int end, int d) {
  uint32_t x = diff_at(e, p, d);
  int len;
  if (s == 0) {
    if (x & 0xffu) return 0;
    if (x) return 1;
    len = 2;
  } else {
    if (x >> 8) return 0;
    len = 1;
  }
  for (int q = p + 1; q < end; ++q) {
    x = (uint32_t)e[padded(q)] ^ (uint32_t)e[padded(q - d)];
    if (x) {
      len += (x & 0xffu) ? 0 : 1;
      break;
    }
    len += 2;
  }
  return len;
}

// Walks a thread's 256 bytes greedily and hands every token to the sink: the longest match
// among the candidate distances when it covers at least three bytes (ties: the nearest),
// else one literal byte.  A match stays inside the thread's chunk; its source may lie
// anywhere earlier in the block.
//   sink.lit(byte);  sink.match(length in bytes, distance in bytes)
// The XORs of the current entry (xp) and the next (xn) against every candidate stay in
// registers: a match of three bytes from the low byte needs xp == 0 and equal low bytes in
// xn, from the high byte an equal high byte in xp and xn == 0 -- so the common case, no
// match, is decided without a branch per candidate, and each entry is loaded once.
template <int MODE, typename Sink>
__device__ __forceinline__ void tokenise(const uint16_t* e, Sink& sink) {
  using C = Cands<MODE>;
  constexpr int K = C::kN;
  const int begin = threadIdx.x * kPerThread, end = begin + kPerThread;
  uint32_t xp[K], xn[K];
  auto load = [&](int q, uint32_t* x) {
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = q < end ? diff_at(e, q, C::at(k)) : kNoMatch;
  };
  int p = begin, s = 0;
  load(p, xp);
  load(p + 1, xn);
  while (p < end) {
    // which candidates give at least three bytes from here
    uint32_t mask = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool ok = s == 0 ? (xp[k] == 0 && (xn[k] & 0xffu) == 0)
                             : ((xp[k] >> 8) == 0 && xn[k] == 0);
      mask |= (uint32_t)ok << k;
    }
    if (mask == 0) {
      const uint32_t v = e[padded(p)];
      sink.lit(s ? v >> 8 : v & 0xffu);
      if (s == 0) {
        s = 1;
      } else {
        s = 0;
        ++p;
#pragma unroll
        for (int k = 0; k < K; ++k) xp[k] = xn[k];
        load(p + 1, xn);
      }
      continue;
    }
    // the longest of them, nearest first; nothing can beat one that reaches the chunk's end
    const int room = 2 * (end - p) - s;
    int best = 0, best_d = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (((mask >> k) & 1u) && best < room) {
        const int len = match_bytes(e, p, s, end, C::at(k));
        if (len > best) { best = len; best_d = C::at(k); }
      }
    }
    sink.match(best, 2 * best_d);
    const int b = 2 * p + s + best;
    p = b >> 1;
    s = b & 1;
    if (p < end) {
      load(p, xp);
      load(p + 1, xn);
    }
  }
}

// ---- survey: CRC-32 of every block, symbol histogram over a sample of the blocks --------
struct HistSink {
  uint32_t* hist;
  __device__ __forceinline__ void lit(uint32_t byte) { atomicAdd(&hist[byte], 1u); }
  __device__ __forceinline__ void match(int len, int dist) {
    int s, x, n;
    length_code(len, s, x, n);
    atomicAdd(&hist[s], 1u);
    distance_code(dist, s, x, n);
    atomicAdd(&hist[286 + s], 1u);
  }
};

__device__ __forceinline__ uint32_t gf2_times(const uint32_t* __restrict__ op, uint32_t v) {
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 32; ++k) out ^= ((v >> k) & 1u) ? op[k] : 0u;
  return out;
}

template <int MODE>
__global__ void __launch_bounds__(kDThreads)
deflate_survey_kernel(const uint16_t* __restrict__ table, uint32_t* __restrict__ hist, uint32_t* __restrict__ block_crc,
                      const uint32_t* __restrict__ crc_ops, int sample_every) {
  __shared__ __align__(16) uint16_t e[kPaddedEntries];
  __shared__ uint32_t crc_table[256];
  __shared__ uint32_t ops[7][32];
  __shared__ uint32_t part[kDThreads];
  __shared__ uint32_t h[kHist];
  const int t = threadIdx.x;
  const uint32_t block = blockIdx.x;
  for (int n = t; n < 256; n += kDThreads) {
    uint32_t c = (uint32_t)n;
#pragma unroll
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0xedb88320u : c >> 1;
    crc_table[n] = c;
  }
  for (int k = t; k < 7 * 32; k += kDThreads) (&ops[0][0])[k] = crc_ops[k];
  for (int k = t; k < kHist; k += kDThreads) h[k] = 0;
  load_block(table, block, e);
  __syncthreads();
  // CRC-32 of the thread's 256 bytes, then a combination tree over the block: level k joins
  // neighbours of 256 << k bytes with the operator for that many zero bytes (crc32_combine)
  uint32_t crc = 0xffffffffu;
  {
    const int begin = t * kPerThread;
#pragma unroll 4
    for (int k = 0; k < kPerThread; ++k) {
      const uint32_t v = e[padded(begin + k)];
      crc = crc_table[(crc ^ v) & 0xffu] ^ (crc >> 8);
      crc = crc_table[(crc ^ (v >> 8)) & 0xffu] ^ (crc >> 8);
    }
    crc ^= 0xffffffffu;
  }
  part[t] = crc;
  __syncthreads();
  for (int level = 0; level < 7; ++level) {
    const int stride = 1 << level;
    uint32_t joined = 0;
    const bool active = (t & (2 * stride - 1)) == 0;
    if (active) joined = gf2_times(ops[level], part[t]) ^ part[t + stride];
    __syncthreads();
    if (active) part[t] = joined;
    __syncthreads();
  }
  if (t == 0) block_crc[block] = part[0];
  if (sample_every > 0 && block % (uint32_t)sample_every == 0) {
    const BlockGeom g = geom<MODE>(block);
    HistSink sink{h};
    tokenise<MODE>(e, sink);
    __syncthreads();
    uint32_t* out = hist + (size_t)g.o * kHist;
    for (int k = t; k < kHist; k += kDThreads)
      if (h[k]) atomicAdd(&out[k], h[k]);
  }
}

// ---- encode -----------------------------------------------------------------------------
constexpr int kThreadWords = 72;   // a thread's private bit buffer: 9 bits per byte on average

// Huffman-codes a thread's tokens into its private bit buffer (word w of thread t at
// buf[w * kDThreads + t]: neighbours in neighbouring banks).  A thread that would overflow
// stops storing and keeps counting; its block is then stored instead.
struct EmitSink {
  const uint32_t* code;   // [316]: reversed code | length << 16
  uint32_t* buf;          // + threadIdx.x
  uint64_t acc;           // pending bits, LSB first
  int nacc;               // number of pending bits (< 32 after a put)
  int word;               // words stored so far
  __device__ __forceinline__ void put(uint32_t v, int n) {
    acc |= (uint64_t)v << nacc;
    nacc += n;
    if (nacc >= 32) {
      if (word < kThreadWords) buf[word * kDThreads] = (uint32_t)acc;
      ++word;
      acc >>= 32;
      nacc -= 32;
    }
  }
  __device__ __forceinline__ void lit(uint32_t byte) {
    const uint32_t a = code[byte];
    put(a & 0xffffu, (int)(a >> 16));
  }
  __device__ __forceinline__ void match(int len, int dist) {
    int s, x, n;
    length_code(len, s, x, n);
    uint32_t c = code[s];
    put((c & 0xffffu) | ((uint32_t)x << (c >> 16)), (int)(c >> 16) + n);
    distance_code(dist, s, x, n);
    c = code[286 + s];
    put((c & 0xffffu) | ((uint32_t)x << (c >> 16)), (int)(c >> 16) + n);
  }
  __device__ __forceinline__ uint32_t finish() {        // -> bits written
    const uint32_t bits = 32u * (uint32_t)word + (uint32_t)nacc;
    if (nacc > 0) {
      if (word < kThreadWords) buf[word * kDThreads] = (uint32_t)acc;
      ++word;
    }
    return bits;
  }
};

template <int MODE>
__global__ void __launch_bounds__(kDThreads)
deflate_encode_kernel(const uint16_t* __restrict__ table,
                      const uint32_t* __restrict__ codes, uint8_t* __restrict__ scratch,
                      uint32_t* __restrict__ sizes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // the entries, later (once every thread has parsed its share) the block's bit stream
  uint16_t* e = reinterpret_cast<uint16_t*>(smem_raw);
  uint32_t* out = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* tbuf = reinterpret_cast<uint32_t*>(smem_raw + kStride);   // kThreadWords x kDThreads
  uint32_t* code = tbuf + kThreadWords * kDThreads;                   // kCodeWords
  uint32_t* scan = code + kCodeWords;                                 // warps + overflow flag
  const int t = threadIdx.x;
  const uint32_t block = blockIdx.x;
  const BlockGeom g = geom<MODE>(block);
  for (int k = t; k < kCodeWords; k += kDThreads) code[k] = codes[(size_t)g.o * kCodeWords + k];
  if (t == 0) scan[kDThreads / 32] = 0;
  load_block(table, block, e);
  __syncthreads();
  EmitSink emit{code, tbuf + t, 0ull, 0, 0};
  tokenise<MODE>(e, emit);
  const uint32_t eob = code[256];
  if (t == kDThreads - 1) {
    emit.put(eob & 0xffffu, (int)(eob >> 16));   // end of block
    emit.put(0u, 3);                              // BFINAL=0, BTYPE=00: empty stored block
  }
  const uint32_t my_bits = emit.finish();
  if (emit.word > kThreadWords) scan[kDThreads / 32] = 1;
  // exclusive scan over the 128 threads (4 warps)
  uint32_t incl = my_bits;
  const int lane = t & 31, warp = t >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) scan[warp] = incl;
  __syncthreads();                 // also: every thread is done with the entries
  uint32_t before = incl - my_bits;
  uint32_t total = 0;
#pragma unroll
  for (int w = 0; w < kDThreads / 32; ++w) {
    if (w < warp) before += scan[w];
    total += scan[w];
  }
  const uint32_t header_bits = code[316];
  // bits of the dynamic block and of the 3-bit header of the empty stored block after it
  const uint32_t body_bits = header_bits + total;
  const uint32_t n_bytes = (body_bits + 7) / 8 + 4;
  uint8_t* dst = scratch + (size_t)block * kStride;
  if (n_bytes > (uint32_t)kBlockBytes || scan[kDThreads / 32] != 0) {
    // would not shrink: a stored block (BFINAL=0, BTYPE=00, LEN, ~LEN, raw bytes)
    if (t == 0) {
      dst[0] = 0;
      dst[1] = (uint8_t)(kBlockBytes & 0xff);
      dst[2] = (uint8_t)(kBlockBytes >> 8);
      dst[3] = (uint8_t)(~kBlockBytes & 0xff);
      dst[4] = (uint8_t)((~kBlockBytes >> 8) & 0xff);
      sizes[block] = 5 + kBlockBytes;
    }
    for (int k = t; k < kBlockEntries; k += kDThreads) {
      const uint16_t v = e[padded(k)];
      dst[5 + 2 * k] = (uint8_t)(v & 0xff);
      dst[6 + 2 * k] = (uint8_t)(v >> 8);
    }
    return;
  }
  for (int k = t; k < kOutWords; k += kDThreads) out[k] = 0;
  __syncthreads();
  // header words, then every thread's bits shifted to its bit offset
  for (uint32_t k = t; k < (header_bits + 31) / 32; k += kDThreads) atomicOr(&out[k], code[317 + k]);
  {
    const uint32_t at = header_bits + before;
    const uint32_t sh = at & 31u;
    uint32_t* o = out + (at >> 5);
    const int n_words = (int)((my_bits + 31) / 32);
    for (int w = 0; w < n_words; ++w) {
      const uint32_t v = tbuf[w * kDThreads + t];
      atomicOr(&o[w], v << sh);
      if (sh) atomicOr(&o[w + 1], v >> (32 - sh));
    }
  }
  __syncthreads();
  if (t == 0) {
    // LEN = 0, NLEN = 0xffff behind the padding bits
    const uint32_t at = (body_bits + 7) / 8;
    uint8_t* ob = reinterpret_cast<uint8_t*>(out);
    ob[at] = 0; ob[at + 1] = 0; ob[at + 2] = 0xff; ob[at + 3] = 0xff;
    sizes[block] = n_bytes;
  }
  __syncthreads();
  const uint4* src4 = reinterpret_cast<const uint4*>(out);
  uint4* dst4 = reinterpret_cast<uint4*>(dst);
  for (uint32_t k = t; k < (n_bytes + 15) / 16; k += kDThreads) dst4[k] = src4[k];
}

// Packs the blocks' bytes back to back: block b goes to out + offsets[b].
__global__ void __launch_bounds__(256)
deflate_gather_kernel(const uint8_t* __restrict__ scratch, const uint32_t* __restrict__ sizes,
                      const int64_t* __restrict__ offsets, uint8_t* __restrict__ out) {
  const uint32_t block = blockIdx.x;
  const uint32_t n = sizes[block];
  const uint8_t* src = scratch + (size_t)block * kStride;
  uint8_t* dst = out + offsets[block];
  // head bytes until the destination is 4-byte aligned, then words assembled from the
  // (16-byte aligned) source with a byte shift, then the tail
  const uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3));
  if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
  const uint32_t n_words = (n - head) / 4;
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + head);
  const uint32_t sh = 8 * head;
  for (uint32_t k = threadIdx.x; k < n_words; k += 256) {
    // bytes head + 4k .. head + 4k + 3 of the source
    const uint32_t lo = s32[k], hi = s32[k + 1];
    d32[k] = sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
  }
  const uint32_t done = head + 4 * n_words;
  if (threadIdx.x < n - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
}

static_assert(kPaddedEntries * 2 <= kStride, "the bit stream reuses the entries' memory");
constexpr size_t kEncodeSmem =
    (size_t)kStride + (size_t)kThreadWords * kDThreads * 4 + (size_t)kCodeWords * 4 + 64 * 4;

}  // namespace
}  // namespace iiv

using namespace iiv;

extern "C" size_t iiv_deflate_block_bytes(void) { return kBlockBytes; }
extern "C" size_t iiv_deflate_block_stride(void) { return kStride; }

static int n_blocks_of(int mode) {
  const int bits = mode == IIV_MODE_HGR ? 14 : 13, n_off = mode == IIV_MODE_HGR ? 2 : 4;
  return (int)(((size_t)n_off << (2 * bits)) / kBlockEntries);
}

extern "C" int iiv_deflate_survey(int mode, const uint16_t* d_table, uint32_t* d_hist, uint32_t* d_block_crc,
                                  const uint32_t* d_crc_ops, int sample_every, void* stream) {
  IIV_REQUIRE(mode == IIV_MODE_HGR || mode == IIV_MODE_DHGR, "bad mode %d", mode);
  IIV_REQUIRE(d_table && d_hist && d_block_crc && d_crc_ops && sample_every >= 0,
              "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_off = mode == IIV_MODE_HGR ? 2 : 4;
  IIV_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * kHist * n_off, st));
  const int n = n_blocks_of(mode);
  if (mode == IIV_MODE_HGR)
    deflate_survey_kernel<IIV_MODE_HGR><<<n, kDThreads, 0, st>>>(d_table, d_hist,
                                                                  d_block_crc, d_crc_ops,
                                                                  sample_every);
  else
    deflate_survey_kernel<IIV_MODE_DHGR><<<n, kDThreads, 0, st>>>(d_table, d_hist,
                                                                   d_block_crc, d_crc_ops,
                                                                   sample_every);
  IIV_LAUNCH_CHECK("deflate_survey_kernel");
  return 0;
}

extern "C" int iiv_deflate_encode(int mode, const uint16_t* d_table,
                                  const uint32_t* d_codes, uint8_t* d_scratch,
                                  uint32_t* d_sizes, void* stream) {
  IIV_REQUIRE(mode == IIV_MODE_HGR || mode == IIV_MODE_DHGR, "bad mode %d", mode);
  IIV_REQUIRE(d_table && d_codes && d_scratch && d_sizes, "null pointer");
  IIV_REQUIRE(((uintptr_t)d_scratch & 15) == 0, "scratch must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = n_blocks_of(mode);
  cudaError_t e;
  if (mode == IIV_MODE_HGR) {
    e = cudaFuncSetAttribute(deflate_encode_kernel<IIV_MODE_HGR>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEncodeSmem);
    if (e == cudaSuccess)
      deflate_encode_kernel<IIV_MODE_HGR><<<n, kDThreads, kEncodeSmem, st>>>(
          d_table, d_codes, d_scratch, d_sizes);
  } else {
    e = cudaFuncSetAttribute(deflate_encode_kernel<IIV_MODE_DHGR>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEncodeSmem);
    if (e == cudaSuccess)
      deflate_encode_kernel<IIV_MODE_DHGR><<<n, kDThreads, kEncodeSmem, st>>>(
          d_table, d_codes, d_scratch, d_sizes);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "deflate_encode_kernel");
  return 0;
}

extern "C" int iiv_deflate_gather(const uint8_t* d_scratch, const uint32_t* d_sizes,
                                  const int64_t* d_offsets, int n_blocks, uint8_t* d_out,
                                  void* stream) {
  IIV_REQUIRE(d_scratch && d_sizes && d_offsets && d_out && n_blocks >= 0, "bad argument");
  if (n_blocks == 0) return 0;
  deflate_gather_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>(d_scratch, d_sizes,
                                                                    d_offsets, d_out);
  IIV_LAUNCH_CHECK("deflate_gather_kernel");
  return 0;
}

// Path 2b: the greedy prioritised delta coder of transcoder/video.py
// (Video.encode_frame :72-93, _index_changes :95-251, _heapify_priorities
//  :253-271, _compute_error :275-301) -- one thread block per clip.
//
// The opcode loop is a sequential dependency chain (each opcode's nonces start where the
// previous one's ended, and its stores change the priorities the next pop reads), so
// one clip cannot use more than one block; the chip is filled by batching independent
// clips.  Inside a block the segment runs in two phases: phase A (all 256 threads)
// scores the bank, draws the heap nonces, radix-selects the reachable prefix of the
// heap and sorts it; phase B cuts the loop into stages run by specialised warps --
// row producers, speculative front ends, one decision warp, a helper (applier + MT19937
// twister) and a preparer of the next segment -- that talk through released / acquired
// words and rings in shared memory (see phase B below).
//
// Exactness notes:
//  * heapq pops the smallest (-priority, nonce, page, offset) tuple; with unique
//    tuples the pop order is the sorted order, so the initial heap is a sorted
//    array consumed by a cursor.  Identical tuples behave identically.
//  * cells re-queued at video.py:177-178 carry -p computed on np.uint16, i.e.
//    65536-p > 0, so they sort after EVERY first-pass cell.  They live in an
//    append-only list that is only searched (block-wide arg-min) once the sorted
//    array is exhausted.
//  * stream N = numpy's global MT19937: np.random.randint(0, 256, size=n) takes
//    the low 8 bits of n successive words, k-th word -> k-th nonzero cell in
//    row-major order (video.py:259-267).  Stream P = CPython's random module:
//    getrandbits(8) = top 8 bits of one word (video.py:178, :291).  Both are
//    advanced on the device from the 624-word states in the clip state blob.
//  * per opcode, every candidate offset (delta < 0) draws a nonce in ascending
//    offset order before any candidate is examined; then each accepted offset
//    with a non-zero residual draws one more (video.py:290-293, :173-178).
//  * at most two further offsets are accepted (len(offsets) == 3 -> break,
//    video.py:181); the fourth slot repeats the first.
#include <mutex>

#include "iiv_common.cuh"
#ifdef IIV_X_TIMING
#include <cstdio>
#endif

namespace iiv {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kProducers = 3;            // row-scoring warps (each keeps two rows in flight)
constexpr int kFronts = 2;               // front-end warps of the opcode loop
// popped entries digested ahead of the decision warp: enough for the front ends to stay busy,
// no more -- the further ahead they read a page, the more often it has changed by the time
// the record is used (8 measured 1-2 % slower, 2 starves the decision warp)
constexpr int kRecRing = 4;
constexpr int kOpQueue = 64;            // emitted opcodes waiting for their stores
// prefetched delta rows in flight: a multiple of the 2 * kProducers entries the producers
// take per round, so that the entries that share a ring slot (e, e + kRing, ...) all belong
// to the same producer warp, which writes them in program order
constexpr int kRing = 18;
constexpr int kPyBlocks = 4;            // resident 624-word blocks of stream P (power of 2)
constexpr int kCells = 32 * 256;       // one bank: 32 pages x 256 offsets
constexpr int kCols = 32 * 128;
constexpr int kPushedCap = 4096;       // re-queued cells kept in shared memory; a segment of
                                       // more than kPushedCap / 2 opcodes may overflow into
                                       // a per-clip scratch list in global memory
constexpr int kMaxBudget = 1 << 17;    // opcodes per segment (one encode_frame generator)
constexpr uint64_t kDead = ~0ull;

// Offsets of the fields inside a clip state blob (bytes).
constexpr size_t kOffPacked = 0;
constexpr size_t kOffMain = kOffPacked + kCols * 8;
constexpr size_t kOffAux = kOffMain + kCells;
constexpr size_t kOffPrioMain = kOffAux + kCells;
constexpr size_t kOffPrioAux = kOffPrioMain + kCells * 4;
constexpr size_t kOffMtNp = kOffPrioAux + kCells * 4;
constexpr size_t kOffMtPy = kOffMtNp + 640 * 4;
constexpr size_t kOffFlags = kOffMtPy + 640 * 4;
constexpr size_t kStateBytes = kOffFlags + 64;

struct Smem {
  uint64_t keys[kCells];       // sorted first-pass heap
  uint64_t src[kCols];         // Video.pixelmap.packed
  uint64_t pushed[kPushedCap]; // re-queued cells (video.py:177-178)
  int32_t prio[kCells];        // update_priority of the active bank
  uint16_t dw[kCells];         // local diff_weights (video.py:109-111)
  uint32_t mt_np[2][624];      // stream N, ping-pong
  uint32_t mt_py[kPyBlocks][624];   // stream P: current block and its successors (ring)
  uint64_t wmin64[8];
  int32_t scan[kThreads / 32];
  uint32_t hist[kThreads / 32][257];   // per-warp digit histograms of the heap select
  uint64_t sel_prefix;
  int sel_remaining, sel_count, sel_done, sel_expect;
  // phase B: rows of new diffs (compute_delta_page's new_diff, video.py:281) for
  // upcoming heap entries, filled by the producer warps.  Slot kRing is the
  // consumer's own (re-queued cells are scored on demand).
  alignas(16) uint16_t ring_row[kRing + 1][256];
  uint32_t ring_tag[kRing];       // sorted-array index the slot holds
  // top byte of the tempered words of stream P (= getrandbits(8), video.py:178, :291):
  // slot s < kPyBlocks holds the block whose number is s (mod kPyBlocks), the last slot
  // repeats slot 0, so that the bytes of the current block and of its successor are
  // always contiguous
  uint8_t py_nonce[(kPyBlocks + 1) * 624 + 16];
  // emitted records waiting for the applier warp to apply their stores; byte 7 of a record
  // carries (sequence number & 255), so one 64-bit store publishes it
  unsigned long long opq[kOpQueue];
  // front-end records: a popped heap entry with its candidate analysis (see phase B)
  // contenders of a record: the competing candidates (delta < 0, priority != 0) whose
  // delta is one of the two smallest -- only they can be among the two winners, whatever
  // the nonces turn out to be.  (delta + 32768) << 16 | nonce rank << 8 | offset.
  uint32_t rec_cont[kRecRing][32];
  // One 16-byte word per record, written with a single store and polled with a single
  // load, so that it is its own "ready" flag (no flag/data ordering to get wrong):
  //   x = entry | kind << 14 | ((record number + 1) & 0xffff) << 16
  //   y = cell | content << 13 | n_cand << 21            (bits 30, 31 are zero)
  //   z = o1 | o2 << 8 | min(n_contenders, 63) << 16 | (r - b_done seen) << 22
  //       | settled << 26 | has1 << 27 | has2 << 28
  //   w = new diff at o1 | new diff at o2 << 16
  // settled: the winners do not depend on the nonces and the front end has already looked
  // them up (o1/o2/has1/has2/w); otherwise the contenders are in rec_cont.
  alignas(16) uint32_t rec[kRecRing][4];
  // (next record number to be popped) << 14 | sorted-heap cursor of that pop: one word, so
  // that turn and cursor cannot be seen out of step
  // (the words below are shared between warps without a barrier: ld_* / st_* helpers only)
  uint32_t pop_state;
  int b_done;            // records the decision warp has finished with
  int final_emitted;     // total records of the segment (set before stop)
  int applied_pub;       // records applied so far
  int head;              // heap entries the consumer is done with
  int stop;              // segment finished: producers leave
  int mt_req, mt_done;   // stream P block twists requested / finished
  int np_pre;            // successor blocks of stream N prepared during phase B
};

__device__ __forceinline__ void twist(const uint32_t* __restrict__ s,
                                      uint32_t* __restrict__ d) {
  const int t = threadIdx.x;
  if (t < 227) d[t] = mt_mix(s[t], s[t + 1], s[t + 397]);
  __syncthreads();
  if (t < 227) d[227 + t] = mt_mix(s[227 + t], s[228 + t], d[t]);
  __syncthreads();
  if (t < 169) d[454 + t] = mt_mix(s[454 + t], s[455 + t], d[227 + t]);
  if (t == 255) d[623] = mt_mix(s[623], d[0], d[396]);
  __syncthreads();
}

// The same generation step done by a single warp (phase B helper warp).
// Also stores getrandbits(8) = top byte of each tempered new word into the block's
// slot(s) of py_nonce (slot s holds block numbers s mod kPyBlocks; one more repeats slot 0).
template <bool kNonces>
__device__ __forceinline__ void warp_twist(const uint32_t* __restrict__ s,
                                           uint32_t* __restrict__ d, int lane, int slot,
                                           uint8_t* __restrict__ py_nonce) {
  uint8_t* nb = py_nonce + slot * 624;
  uint8_t* nb2 = py_nonce + (slot == 0 ? kPyBlocks * 624 : slot * 624);
  // three dependent sweeps of 227 / 227 / 170 words; inside a sweep every word is
  // independent.  Unrolled by two only: the opcode loop's roles together are several times
  // the instruction cache, and this one has time to spare
#pragma unroll 2
  for (int k = 0; k < 8; ++k) {
    const int t = lane + 32 * k;
    if (t < 227) {
      const uint32_t w = mt_mix(s[t], s[t + 1], s[t + 397]);
      d[t] = w;
      if (kNonces) nb[t] = nb2[t] = (uint8_t)(mt_temper(w) >> 24);
    }
  }
  __syncwarp();
#pragma unroll 2
  for (int k = 0; k < 8; ++k) {
    const int t = lane + 32 * k;
    if (t < 227) {
      const uint32_t w = mt_mix(s[227 + t], s[228 + t], d[t]);
      d[227 + t] = w;
      if (kNonces) nb[227 + t] = nb2[227 + t] = (uint8_t)(mt_temper(w) >> 24);
    }
  }
  __syncwarp();
#pragma unroll 2
  for (int k = 0; k < 6; ++k) {
    const int t = lane + 32 * k;
    if (t < 169) {
      const uint32_t w = mt_mix(s[454 + t], s[455 + t], d[227 + t]);
      d[454 + t] = w;
      if (kNonces) nb[454 + t] = nb2[454 + t] = (uint8_t)(mt_temper(w) >> 24);
    }
  }
  if (lane == 0) {
    const uint32_t w = mt_mix(s[623], d[0], d[396]);
    d[623] = w;
    if (kNonces) nb[623] = nb2[623] = (uint8_t)(mt_temper(w) >> 24);
  }
  __syncwarp();
}

// ---- cross-warp hand-offs of phase B, on the PTX memory model --------------------------
// Every word that one warp writes while another may read it is accessed with a strong
// (.relaxed / .acquire / .release, scope .cta) operation: a flag is published with a release
// and polled with an acquire, and what sits behind it may then be read with plain loads.
// On sm_100a an acquire or relaxed load of shared memory is a plain LDS and a relaxed store
// a plain STS -- the qualifiers only bind ptxas, which may otherwise hoist a load above the
// spin loop that guards it (it has) -- while a release is MEMBAR.ALL.CTA + STS.
__device__ __forceinline__ uint32_t smem_addr(const volatile void* p) {
  return (uint32_t)__cvta_generic_to_shared(const_cast<const void*>(p));
}
__device__ __forceinline__ uint32_t ld_acq_u32(const volatile void* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_rlx_u32(const volatile void* p) {
  uint32_t v;
  asm volatile("ld.relaxed.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_addr(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_rlx_u16(const volatile void* p) {
  uint16_t v;
  asm volatile("ld.relaxed.cta.shared.u16 %0, [%1];" : "=h"(v) : "r"(smem_addr(p)) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_rlx_u64(const volatile void* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.cta.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_addr(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_acq_v4(const volatile void* p) {
  uint4 v;
  asm volatile("ld.acquire.cta.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_rlx_v4(const volatile void* p) {
  uint4 v;
  asm volatile("ld.relaxed.cta.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_rel_u32(volatile void* p, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_addr(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_rlx_u32(volatile void* p, uint32_t v) {
  asm volatile("st.relaxed.cta.shared.u32 [%0], %1;" ::"r"(smem_addr(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_rlx_u16(volatile void* p, uint16_t v) {
  asm volatile("st.relaxed.cta.shared.u16 [%0], %1;" ::"r"(smem_addr(p)), "h"(v) : "memory");
}
__device__ __forceinline__ void st_rlx_u64(volatile void* p, unsigned long long v) {
  asm volatile("st.relaxed.cta.shared.u64 [%0], %1;" ::"r"(smem_addr(p)), "l"(v) : "memory");
}
__device__ __forceinline__ void st_rel_v4(volatile void* p, uint4 v) {
  asm volatile("st.release.cta.shared.v4.u32 [%0], {%1, %2, %3, %4};"
               ::"r"(smem_addr(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_rlx_v4(volatile void* p, uint4 v) {
  asm volatile("st.relaxed.cta.shared.v4.u32 [%0], {%1, %2, %3, %4};"
               ::"r"(smem_addr(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// fence.acq_rel.cta (MEMBAR.ALL.CTA): with a relaxed store after it, a release pattern on
// every word stored later in program order (PTX memory model, release patterns, case 3)
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
// a slot lock of the row ring: acquire on taking it, release on giving it back
__device__ __forceinline__ uint32_t cas_acq_u32(volatile void* p, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.acquire.cta.shared.cas.b32 %0, [%1], %2, %3;"
               : "=r"(old) : "r"(smem_addr(p)), "r"(cmp), "r"(val) : "memory");
  return old;
}

// getrandbits(8) bytes of one block of stream P into its slot(s) of py_nonce.
__device__ __forceinline__ void fill_nonces(const uint32_t* __restrict__ block, int slot,
                                            uint8_t* __restrict__ py_nonce, int idx,
                                            int stride) {
#pragma unroll 4
  for (int k = idx; k < 624; k += stride) {
    const uint8_t b = (uint8_t)(mt_temper(block[k]) >> 24);
    py_nonce[slot * 624 + k] = b;
    if (slot == 0) py_nonce[kPyBlocks * 624 + k] = b;
  }
}

// Row of new diffs for storing `content` at every offset of `page` of the target
// (Bitmap._diff_weights_page with source = target, screen.py:453-494, 544): lane l
// owns offsets 8l .. 8l+7, i.e. packed columns 4l .. 4l+3.  Screen holes (lanes 15
// and 31) get 0xffff so that they are never candidates.
// (the lane's 8 values are returned packed as 16-bit pairs)
template <int MODE>
__device__ __forceinline__ uint4 score_row_regs(const uint64_t* __restrict__ tp_row,
                                                const uint16_t* __restrict__ table,
                                                uint32_t content, int is_aux, int lane) {
  using M = Mode<MODE>;
  uint4 packed_out = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  if ((lane & 15) != 15) {
    const ulonglong2 w01 = __ldg(reinterpret_cast<const ulonglong2*>(tp_row) + 2 * lane);
    const ulonglong2 w23 = __ldg(reinterpret_cast<const ulonglong2*>(tp_row) + 2 * lane + 1);
    const uint64_t w[4] = {w01.x, w01.y, w23.x, w23.y};
    uint32_t nd[8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int o = byte_offset<MODE>(half, is_aux);
        const uint32_t x = mask_shift<MODE>(masked_update<MODE>(o, w[q], content), o);
        const uint32_t y = mask_shift<MODE>(w[q], o);
        nd[2 * q + half] =
            ldg_table(table + ((size_t)o << (2 * M::kBits)) + ((size_t)x << M::kBits) + y);
      }
    packed_out = make_uint4(nd[0] | (nd[1] << 16), nd[2] | (nd[3] << 16),
                            nd[4] | (nd[5] << 16), nd[6] | (nd[7] << 16));
  }
  return packed_out;
}

// Out of line on purpose: its one caller is the decision warp's rare "heap ran dry" path,
// and that warp's loop should stay small (the roles of phase B outgrow the instruction
// cache several times over).
template <int MODE>
__device__ __noinline__ void score_row(const uint64_t* __restrict__ tp_row,
                                          const uint16_t* __restrict__ table,
                                          uint32_t content, int is_aux, int lane,
                                          uint16_t* __restrict__ out_row) {
  reinterpret_cast<uint4*>(out_row)[lane] =
      score_row_regs<MODE>(tp_row, table, content, is_aux, lane);
}

__device__ __forceinline__ uint64_t first_pass_key(int32_t prio, uint32_t nonce,
                                                   int cell) {
  // ascending (-prio, nonce, page, offset); cell = page * 256 + offset.
  return ((uint64_t)(0x7fffffffu - (uint32_t)prio) << 21) | ((uint64_t)nonce << 13) |
         (uint64_t)cell;
}

__device__ __forceinline__ uint64_t requeue_key(uint32_t p, uint32_t nonce, int cell) {
  // ascending (65536 - p, nonce, page, offset).
  return ((uint64_t)((65536u - p) & 0xffffu) << 21) | ((uint64_t)nonce << 13) |
         (uint64_t)cell;
}

__device__ __forceinline__ uint64_t warp_min64(uint64_t v) {
  const uint32_t hi = (uint32_t)(v >> 32);
  const uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
  const uint32_t lo = hi == mh ? (uint32_t)v : 0xffffffffu;
  const uint32_t ml = __reduce_min_sync(0xffffffffu, lo);
  return ((uint64_t)mh << 32) | ml;
}

// Bitonic steps of partner distance <= 32 on one 64-key chunk, for the network phases
// k_lo .. k_hi (powers of two; phases above 64 contribute their last six steps only).
// `base` is the chunk's index in the whole array: the direction of a compare-exchange
// is ascending where (index & k) == 0.  Lane l holds keys l and l + 32 of the chunk.
__device__ __forceinline__ void sort_chunk64(uint64_t* chunk, int base, int lane, int k_lo,
                                             int k_hi) {
  uint64_t e0 = chunk[lane], e1 = chunk[lane + 32];
  for (int k = k_lo; k <= k_hi; k <<= 1) {
    const bool asc0 = ((base + lane) & k) == 0;
    const bool asc1 = ((base + lane + 32) & k) == 0;
    if (k >= 64) {   // distance 32: the lane's own pair (same direction for both)
      if ((e0 > e1) == asc0) {
        const uint64_t x = e0;
        e0 = e1;
        e1 = x;
      }
    }
    for (int j = min(k >> 1, 16); j > 0; j >>= 1) {
      const uint64_t p0 = __shfl_xor_sync(0xffffffffu, e0, j);
      const uint64_t p1 = __shfl_xor_sync(0xffffffffu, e1, j);
      const bool lower = (lane & j) == 0;
      // the lower index keeps the minimum where the direction is ascending
      e0 = ((lower == asc0) == (p0 < e0)) ? p0 : e0;
      e1 = ((lower == asc1) == (p1 < e1)) ? p1 : e1;
    }
  }
  chunk[lane] = e0;
  chunk[lane + 32] = e1;
}

template <int MODE>
__device__ __forceinline__ void apply_store(uint64_t* src, uint8_t* mem, int page,
                                            int offset, int is_aux, uint32_t value) {
  const int o = byte_offset<MODE>(offset, is_aux);
  const int c = offset >> 1;
  // relaxed stores: the helper warp is the only writer in phase B, but the stream-N warp
  // reads these words (for prefetch addresses only) while they change
  uint64_t* row = src + page * 128;
  const uint64_t w = masked_update<MODE>(o, row[c], value);
  st_rlx_u64(&row[c], w);
  if (o == 0 && c > 0)
    st_rlx_u64(&row[c - 1], (row[c - 1] & keep_low_mask<MODE>()) ^ footer_of<MODE>(w));
  else if (o == Mode<MODE>::kOffsets - 1 && c < 127)
    st_rlx_u64(&row[c + 1], (row[c + 1] & keep_high_mask<MODE>()) ^ header_of<MODE>(w));
  mem[page * 256 + offset] = (uint8_t)value;
}

// Optional extras of a launch (all zero for the plain batch call).
struct EncodeExtras {
  int3 inline_segment;      // the schedule of a one-segment launch (segments == nullptr)
  const uint8_t* state_in;  // clip 0 starts from this blob instead of its own (one clip only)
  uint8_t* tail_out;        // the final state's tail (generators + flags) is ALSO written here
  uint8_t* opcodes_mirror;  // every segment's opcode records are ALSO written here (clip 0)
};

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
encode_kernel(uint8_t* __restrict__ states, size_t state_stride,
              const uint8_t* __restrict__ target_mem,
              const uint64_t* __restrict__ target_packed, int n_frames,
              const int32_t* __restrict__ segments, int n_segments,
              const uint16_t* __restrict__ table, uint8_t* __restrict__ opcodes,
              int64_t total_budget, int64_t* __restrict__ seg_info,
              uint64_t* __restrict__ overflow, int overflow_cap,
              const __grid_constant__ EncodeExtras extras) {
  using M = Mode<MODE>;
  constexpr int kBanks = MODE == IIV_MODE_DHGR ? 2 : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int clip = blockIdx.x;
  uint8_t* state = states + (size_t)clip * state_stride;
  const int3 inline_segment = extras.inline_segment;
  if (extras.state_in != nullptr && extras.state_in != state) {
    // start from another blob: everything the kernel does not rewrite must carry over
    const uint4* in = reinterpret_cast<const uint4*>(extras.state_in);
    uint4* out = reinterpret_cast<uint4*>(state);
    for (int k = threadIdx.x; k < (int)(kStateBytes / 16); k += kThreads) out[k] = in[k];
    __syncthreads();
  }
  uint64_t* g_packed = reinterpret_cast<uint64_t*>(state + kOffPacked);
  uint32_t* g_mt_np = reinterpret_cast<uint32_t*>(state + kOffMtNp);
  uint32_t* g_mt_py = reinterpret_cast<uint32_t*>(state + kOffMtPy);
  int32_t* g_flags = reinterpret_cast<int32_t*>(state + kOffFlags);

  // ---- load persistent state ---------------------------------------------------
  for (int k = t; k < kCols; k += kThreads) sm.src[k] = g_packed[k];
  for (int k = t; k < 624; k += kThreads) {
    sm.mt_np[0][k] = g_mt_np[k];
    sm.mt_py[0][k] = g_mt_py[k];
  }
  int np_cur = 0;                 // which ping-pong buffer holds stream N
  // Successor blocks of stream N that the stream-N warp generated during the previous
  // segment's opcode loop; they live in the upper half of the heap array, which phase B
  // does not use when the sorted prefix is at most 4096 keys.
  constexpr int kNpPreMax = 13;   // (623 + 7680 - 1) / 624 blocks at most per segment
  uint32_t* const np_pre_blocks = reinterpret_cast<uint32_t*>(&sm.keys[kCells / 2]);
  static_assert(kNpPreMax * 624 * 4 <= (kCells / 2) * 8, "pre-twisted blocks do not fit");
  int np_ready = 0;
  int pos_np = (int)g_mt_np[624]; // 0..624
  int pos_py = (int)g_mt_py[624];
  // Three blocks of stream P stay resident (ring of buffers, slot = block number mod 3):
  // when the position moves on to the next block the twister refills the freed buffer a
  // whole block (~8 opcodes) before it can be needed.
  int py_cur = 0;                 // slot of the current block
  __syncthreads();
  for (int b = 1; b < kPyBlocks; ++b) twist(sm.mt_py[b - 1], sm.mt_py[b]);
  for (int b = 0; b < kPyBlocks; ++b) fill_nonces(sm.mt_py[b], b, sm.py_nonce, t, kThreads);
  int error_flags = 0;

  uint8_t* op_out = opcodes + (size_t)clip * total_budget * 8;
  // re-queued cells beyond the shared-memory list (only segments of more than
  // kPushedCap / 2 opcodes can get there): written and read by the decision warp alone
  volatile uint64_t* const ovf =
      overflow != nullptr ? overflow + (size_t)clip * (size_t)overflow_cap : nullptr;
  const int pushed_cap = kPushedCap + (ovf != nullptr ? overflow_cap : 0);

  for (int seg = 0; seg < n_segments; ++seg) {
    // (a single segment may ride in the kernel parameters: segments == nullptr)
    const int frame = segments ? segments[3 * seg + 0] : inline_segment.x;
    const int is_aux = segments ? segments[3 * seg + 1] : inline_segment.y;
    const int budget = segments ? segments[3 * seg + 2] : inline_segment.z;
    int64_t* info = seg_info + ((size_t)clip * n_segments + seg) * 8;
    const long long clk_seg = clock64();
    if (budget <= 0) {  // generator created but never pulled: no side effects
      if (t == 0)
        for (int k = 0; k < 8; ++k) info[k] = 0;
      continue;
    }
    const int bank = (MODE == IIV_MODE_DHGR && is_aux) ? 1 : 0;
    const uint64_t* tp = target_packed + ((size_t)clip * n_frames + frame) * kCols;
    const uint8_t* tmem =
        target_mem + (((size_t)clip * n_frames + frame) * kBanks + bank) * kCells;
    int32_t* g_prio = reinterpret_cast<int32_t*>(
        state + (is_aux ? kOffPrioAux : kOffPrioMain));
    uint8_t* g_mem = state + (is_aux ? kOffAux : kOffMain);

#ifdef IIV_X_TIMING
    __shared__ long long pa_sh[8];
    long long pa[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long pa_prev = clock64();
#endif
    // ======================= phase A: score + heapify ===========================
    // thread t owns offset t of every page: cells 256 k + t, k = 0..31.  Lanes thus hold
    // consecutive cells, so every shared-memory access below -- source words, priorities,
    // diff weights, keys -- is to consecutive addresses (32 consecutive cells per thread
    // would put all 32 lanes on one bank: the fold and the key build then cost more than the
    // gathers), and the global loads of target words and priorities are coalesced.
    // All loads of a stage are issued before anything waits on them: the target words
    // first, then the 32 table gathers.
    const int half_a = t & 1;
    const int o_a = byte_offset<MODE>(half_a, is_aux);
    uint32_t dwv[32];
    {
      uint64_t g[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) g[k] = __ldg(tp + 128 * k + (t >> 1));
      uint32_t gidx[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const uint64_t s = sm.src[128 * k + (t >> 1)];
        const uint32_t x = mask_shift<MODE>(s, o_a), y = mask_shift<MODE>(g[k], o_a);
        // a window that already shows what the target shows is at distance 0 (the tables'
        // diagonal): no gather -- on real footage most of the screen; holes likewise
        // (video.py:111: offset t of any page)
        gidx[k] = (x == y || is_hole(t))
                      ? 0xffffffffu
                      : ((uint32_t)o_a << (2 * M::kBits)) + (x << M::kBits) + y;
      }
#pragma unroll
      for (int k = 0; k < 32; ++k)
        dwv[k] = gidx[k] == 0xffffffffu ? 0u : ldg_table(table + gidx[k]);
    }
    // While the gathers are in flight: the draws of stream N that are already at hand -- the
    // current block and the successors the twister prepared during the last opcode loop --
    // go to the per-draw array (more of them than this heapify may need; which is known
    // only once the gathers are back).  The array borrows the row ring, idle until phase B.
    uint8_t* np_nonce = reinterpret_cast<uint8_t*>(&sm.ring_row[0][0]);
    static_assert(sizeof(sm.ring_row) >= kCells, "nonce scratch too small");
    const int np_early = min(np_ready, kNpPreMax);
    for (int b = 0; b <= np_early; ++b) {
      const uint32_t* blk = b == 0 ? sm.mt_np[np_cur] : np_pre_blocks + (b - 1) * 624;
      for (int k = t; k < 624; k += kThreads) {
        const int g = b * 624 + k - pos_np;
        if (g >= 0 && g < kCells) np_nonce[g] = (uint8_t)(mt_temper(blk[k]) & 0xffu);
      }
    }
#ifdef IIV_X_TIMING
    {
      uint32_t chk = 0;
#pragma unroll
      for (int k = 0; k < 32; ++k) chk += dwv[k];
      if (__syncthreads_or(chk == 0xfffffff0u)) pa[7] = 1;
      if (t == 0) { const long long tn = clock64(); pa[6] = tn - pa_prev; pa_prev = tn; }
    }
#endif
    int64_t prio_sum = 0;
    uint32_t nzmask = 0;
    int32_t p_hi = 0, p_lo = 0x7fffffff;      // extremes of the nonzero priorities
    {
      int32_t pv[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) pv[k] = g_prio[256 * k + t];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        int32_t p = pv[k];
        prio_sum += p;
        if (dwv[k] == 0) p = 0;          // video.py:115
        p += (int32_t)dwv[k];            // video.py:116
        sm.prio[256 * k + t] = p;
        sm.dw[256 * k + t] = (uint16_t)dwv[k];
        if (p != 0) {
          nzmask |= 1u << k;
          p_hi = max(p_hi, p);
          p_lo = min(p_lo, p);
        }
      }
    }
    // Row-major rank of every nonzero cell (its draw from stream N, video.py:259-267): page k
    // of warp w holds cells 256 k + 32 w .. + 31, so the counts are scanned in (k, w) order.
    uint32_t* cnt = &sm.hist[0][0];      // [32 pages][8 warps]; the select's scratch, idle now
    static_assert(sizeof(sm.hist) >= 32 * (kThreads / 32) * 4 + 4, "rank scratch too small");
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const uint32_t bal = __ballot_sync(0xffffffffu, (nzmask >> k) & 1u);
      if (lane == k) cnt[k * (kThreads / 32) + warp] = (uint32_t)__popc(bal);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
      prio_sum += __shfl_xor_sync(0xffffffffu, prio_sum, d);
    p_hi = __reduce_max_sync(0xffffffffu, p_hi);
    p_lo = __reduce_min_sync(0xffffffffu, p_lo);
    if (lane == 0) {
      sm.wmin64[warp] = (uint64_t)prio_sum;
      cnt[264 + warp] = (uint32_t)p_hi;
      cnt[272 + warp] = (uint32_t)p_lo;
    }
    __syncthreads();
    if (warp == 0) {
      // exclusive scan of the 256 counts: lane l owns entries 8 l .. 8 l + 7 (page l)
      uint32_t c[8], tot = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        c[q] = cnt[8 * lane + q];
        tot += c[q];
      }
      uint32_t incl = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      uint32_t run = incl - tot;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        cnt[8 * lane + q] = run;
        run += c[q];
      }
      if (lane == 31) cnt[256] = incl;         // n_heap
    }
    __syncthreads();
#ifdef IIV_X_TIMING
    if (t == 0) { const long long tn = clock64(); pa[1] = tn - pa_prev; pa_prev = tn; }
#endif
    const int n_heap = (int)cnt[256];
    int64_t prio_total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) prio_total += (int64_t)sm.wmin64[w];
    // this thread's rank bases, before the scratch is reused: base of (page k, this warp)
    uint32_t rank_base[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) rank_base[k] = cnt[k * (kThreads / 32) + warp];
    // The select below counts the keys by one 8-bit digit at a time, from the highest bit in
    // which any two keys can differ: the keys of the largest and the smallest priority bound
    // them all, so that bit is known before a key exists and the first count is taken while
    // the keys are being built.
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      p_hi = max(p_hi, (int32_t)cnt[264 + w]);
      p_lo = min(p_lo, (int32_t)cnt[272 + w]);
    }
    const uint64_t key_span = first_pass_key(p_hi, 0u, 0) ^ first_pass_key(p_lo, 0xffu, kCells - 1);
    const int top_shift = max(0, 63 - __clzll((long long)key_span) - 7);
    __syncthreads();       // everybody has read the scratch: it becomes the histograms
    for (int k = t; k < (kThreads / 32) * 257; k += kThreads) (&sm.hist[0][0])[k] = 0;
    // stream N draws: the k-th nonzero cell (row-major) takes the low byte of word
    // pos_np + k (video.py:259-267).  Blocks of 624 words are generated one after the
    // other (each needs its predecessor); the bytes of the blocks that were at hand are in
    // the per-draw array already (above).
    const int twists = n_heap > 0 ? (pos_np + n_heap - 1) / 624 : 0;
    {
      const uint32_t* blk = sm.mt_np[np_cur];
      for (int b = 0; b <= twists; ++b) {
        if (b > 0) {
          if (b <= np_ready) {
            blk = np_pre_blocks + (b - 1) * 624;
          } else {
            twist(blk, sm.mt_np[np_cur ^ 1]);
            np_cur ^= 1;
            blk = sm.mt_np[np_cur];
          }
        }
        if (b > np_early)
          for (int k = t; k < 624; k += kThreads) {
            const int g = b * 624 + k - pos_np;
            if (g >= 0 && g < n_heap) np_nonce[g] = (uint8_t)(mt_temper(blk[k]) & 0xffu);
          }
      }
      // the block the position ends in becomes the resident state (before the keys below
      // overwrite the prepared blocks)
      if (blk != sm.mt_np[np_cur])
        for (int k = t; k < 624; k += kThreads) sm.mt_np[np_cur][k] = blk[k];
      np_ready = 0;
    }
    __syncthreads();
#ifdef IIV_X_TIMING
    if (t == 0) { const long long tn = clock64(); pa[2] = tn - pa_prev; pa_prev = tn; }
#endif
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const uint32_t bal = __ballot_sync(0xffffffffu, (nzmask >> k) & 1u);
      if ((nzmask >> k) & 1u) {
        const int r = (int)rank_base[k] + __popc(bal & ((1u << lane) - 1u));
        const uint64_t key = first_pass_key(sm.prio[256 * k + t], np_nonce[r], 256 * k + t);
        sm.keys[r] = key;
        atomicAdd(&sm.hist[warp][(uint32_t)(key >> top_shift) & 255u], 1u);
      }
    }
    __syncthreads();
#ifdef IIV_X_TIMING
    if (t == 0) { const long long tn = clock64(); pa[3] = tn - pa_prev; pa_prev = tn; }
#endif
    if (n_heap > 0) pos_np = pos_np + n_heap - 624 * twists;
    // Only a prefix of the heap can ever be popped in this segment: every opcode
    // pops one live entry and can zero at most two others (video.py:170), so no more
    // than 3 * budget entries are consumed.  Select exactly those (MSB-first radix
    // select of the need-th smallest key; keys are unique) and sort only them.
    int n_sorted = n_heap;
    {
      const int need = min(n_heap, 3 * budget);
      if (need < n_heap && need <= kPushedCap && n_heap > 1024) {
        // (first digit = the eight bits ending at the highest bit the keys can differ in, so
        // that its 256 buckets are in use and the boundary bucket is small: the count taken
        // during the key build usually settles the selection)
        // The sort pads to a power of two anyway, so the select may stop as soon as the
        // whole bucket holding the need-th key fits in that padding: `need` keys are still
        // guaranteed, a few more ride along for free.
        int cap = 256;
        while (cap < need) cap <<= 1;
        if (cap > kPushedCap) cap = kPushedCap;
        if (t == 0) {
          sm.sel_prefix = (sm.keys[0] >> (top_shift + 8)) << (top_shift + 8);
          sm.sel_remaining = need;
          sm.sel_done = 0;
          sm.sel_expect = need;
        }
        __syncthreads();
        // bits >= fixed of the prefix are settled; a digit may overlap them on the last
        // pass (shift clamped to 0), where its upper bits are then equal for every match
        bool counted = true;      // the first digit's histogram was taken with the keys
        for (int shift = top_shift, fixed = top_shift + 8; !sm.sel_done;
             fixed = shift, shift = max(shift - 8, 0)) {
          if (!counted) {
            for (int k = t; k < (kThreads / 32) * 257; k += kThreads) (&sm.hist[0][0])[k] = 0;
            __syncthreads();
          }
          const uint64_t prefix = sm.sel_prefix;
          const int remaining = sm.sel_remaining;
          if (!counted) {
            for (int k = t; k < n_heap; k += kThreads) {
              const uint64_t key = sm.keys[k];
              if (((key ^ prefix) >> fixed) == 0)
                atomicAdd(&sm.hist[warp][(uint32_t)(key >> shift) & 255u], 1u);
            }
            __syncthreads();
          }
          counted = false;
          if (warp == 0) {
            // lane l owns digits 8l .. 8l+7
            uint32_t c[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint32_t v = 0;
#pragma unroll
              for (int w = 0; w < kThreads / 32; ++w) v += sm.hist[w][8 * lane + q];
              c[q] = v;
              tot += v;
            }
            uint32_t incl = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
              if (lane >= d) incl += v;
            }
            const uint32_t excl = incl - tot;
            // the lane whose digit range contains the remaining-th key of this prefix
            if (excl < (uint32_t)remaining && (uint32_t)remaining <= incl) {
              uint32_t run = excl;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (run < (uint32_t)remaining && (uint32_t)remaining <= run + c[q]) {
                  const uint64_t pre = prefix | ((uint64_t)(8 * lane + q) << shift);
                  const int taken = need - remaining + (int)run;   // keys below this bucket
                  if (taken + (int)c[q] <= cap && shift > 0) {
                    sm.sel_prefix = pre | ((1ull << shift) - 1ull);   // the whole bucket
                    sm.sel_expect = taken + (int)c[q];
                    sm.sel_done = 1;
                  } else {
                    sm.sel_prefix = pre;
                    sm.sel_remaining = remaining - (int)run;
                    if (shift == 0) sm.sel_done = 1;   // pre is the need-th key itself
                  }
                }
                run += c[q];
              }
            }
          }
          __syncthreads();
        }
        // compact the keys <= threshold (sel_expect >= need of them) through the re-queue
        // buffer, which is idle until phase B
        const uint64_t threshold = sm.sel_prefix;
        const int n_sel = sm.sel_expect;
        if (t == 0) sm.sel_count = 0;
        __syncthreads();
        for (int k0 = 0; k0 < n_heap; k0 += kThreads) {
          const int k = k0 + t;
          const uint64_t key = k < n_heap ? sm.keys[k] : kDead;
          const bool keep = key <= threshold;
          const uint32_t bal = __ballot_sync(0xffffffffu, keep);
          int basepos = 0;
          if (lane == 0 && bal) basepos = atomicAdd(&sm.sel_count, __popc(bal));
          basepos = __shfl_sync(0xffffffffu, basepos, 0);
          if (keep) sm.pushed[basepos + __popc(bal & ((1u << lane) - 1u))] = key;
        }
        __syncthreads();
        n_sorted = n_sel;
        if (sm.sel_count != n_sel) error_flags |= 2;   // cannot happen: keys are unique
        for (int k = t; k < n_sel; k += kThreads) sm.keys[k] = sm.pushed[k];
        __syncthreads();
      }
    }
    // pad to a power of two and sort ascending.
    int P = 256;
    while (P < n_sorted) P <<= 1;
    for (int k = n_sorted + t; k < P; k += kThreads) sm.keys[k] = kDead;
    __syncthreads();
#ifdef IIV_X_TIMING
    if (t == 0) { const long long tn = clock64(); pa[4] = tn - pa_prev; pa_prev = tn; }
#endif
    // Bitonic network.  Steps whose partner distance is below 64 never leave a 64-key
    // chunk, so a warp runs all of them back to back on a chunk held in registers (two
    // keys per lane, shuffles for distances < 32); only the wider steps go through
    // shared memory with a block barrier: 15 barriers for 1024 keys instead of 55.
    for (int c = warp; c < P / 64; c += kWarps) sort_chunk64(sm.keys + 64 * c, 64 * c, lane, 2, 64);
    __syncthreads();
    for (int k = 128; k <= P; k <<= 1) {
      for (int j = k >> 1; j >= 64; j >>= 1) {
        for (int idx = t; idx < P / 2; idx += kThreads) {
          const int i = ((idx & ~(j - 1)) << 1) | (idx & (j - 1));
          const int l = i | j;
          const uint64_t a = sm.keys[i], b = sm.keys[l];
          const bool asc = (i & k) == 0;
          if ((a > b) == asc) {
            sm.keys[i] = b;
            sm.keys[l] = a;
          }
        }
        __syncthreads();
      }
      for (int c = warp; c < P / 64; c += kWarps) sort_chunk64(sm.keys + 64 * c, 64 * c, lane, k, k);
      __syncthreads();
    }
#ifdef IIV_X_TIMING
    if (t == 0) { const long long tn = clock64(); pa[5] = tn - pa_prev; pa_prev = tn; }
    if (t == 0) for (int k = 0; k < 8; ++k) pa_sh[k] = pa[k];
#endif
    const int n_first = n_sorted;   // entries of the sorted array phase B may walk

    // ======================= phase B: emit opcodes ==============================
    // The opcode loop is a chain -- each opcode's nonces start where the previous
    // one's ended, and its stores change the priorities the next pop sees -- so it is
    // cut into stages run by specialised warps, none of which crosses a block barrier:
    //   producers (warps 0-2)  run ahead along the sorted heap and score the row of
    //       new diffs of each upcoming entry (two dependent global loads: target word,
    //       table gather) into a ring in shared memory, two entries in flight per warp.
    //       A row depends on the target frame and the content byte only, never on the
    //       evolving source, so it cannot go stale.
    //   front ends (warps 5, 6)  take turns popping the next live heap entry and
    //       digest it SPECULATIVELY against the current priorities / diff weights of
    //       its page: candidate bits, eligibility bits, nonce ranks and key prefixes of
    //       all 256 offsets go into a record ring.
    //   decision (warp 7)  consumes the records in order.  A record is valid unless an
    //       opcode decided after the front end read the page touched the same page
    //       (only stores to that page change what was read); then the warp re-digests
    //       the entry itself.  It draws the nonces, picks the two best offsets, updates
    //       priorities, re-queues, publishes the opcode: shared memory + shuffles only.
    //   helper (warp 4)  two background jobs: it commits the stores of every published
    //       opcode to the source bitmap and memory map (Bitmap.apply), in order -- nothing
    //       in phase B reads the source, so this is off the chain -- and it prepares the
    //       next MT19937 block of stream P when asked.  (One warp instead of two: a polling
    //       warp costs its scheduler's other warps issue slots -- 148 clips +9 %, a single
    //       clip +2 %.)
    //   np (warp 3)  makes the stream N blocks the next segment's heapify will draw from,
    //       then goes quiet.
    // A cell whose priority is already 0 is skipped by everyone alike: priorities only
    // ever fall to 0 inside a segment (video.py:140, :159-170).
    // The issue arbiter favours the highest warp id of a scheduler and a spinning warp
    // is nearly always eligible, hence the decision warp is 7 and shares its scheduler
    // (warp id % 4) only with the np warp, which works for a tenth of a segment and never
    // polls (with the helper on that scheduler instead a clip is 2.4 % slower); helper loops
    // back off with nanosleep.
    constexpr uint32_t kFull = 0xffffffffu;
        constexpr int kDecideWarp = 7, kHelpWarp = 4, kNpWarp = 3;
    constexpr uint32_t kKindLive = 0u, kKindDead = 1u, kKindEndOfHeap = 2u;
    if (t < kRing) {
      sm.ring_tag[t] = 0xffffffffu;
    }
    if (t < kRecRing) *reinterpret_cast<uint4*>(sm.rec[t]) = make_uint4(0u, 0u, 0u, 0u);
    if (t == 0) {
      sm.head = 0;
      sm.stop = 0;
      sm.mt_req = 0;
      sm.mt_done = 0;
      sm.np_pre = 0;
      sm.final_emitted = 0;
      sm.applied_pub = 0;
      sm.pop_state = 0u;
      sm.b_done = 0;
    }
    if (t < kOpQueue) sm.opq[t] = 0ull;
    __syncthreads();
    uint8_t* seg_out = op_out;
    int emitted = 0, py_words = 0;
    bool out_of_work = false;
    const long long clk_b = clock64();
    long long wait_rows = 0, wait_misc = 0, wait_mt = 0;   // decision-warp stall cycles (diagnostics)
    uint32_t* tags = sm.ring_tag;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // Candidate analysis of the 8 offsets a lane owns on `page` for the entry at
    // `cell` whose row of new diffs is `row`: candidate bits (delta < 0, video.py:283),
    // eligibility bits (priority != 0, video.py:159), nonce rank of the lane's first
    // candidate, key prefixes, and the page's candidate count.
    // The caller has acquired the row (its tag) and, in a front end, sampled b_done before:
    // the page's words race with the decision warp's stores by design -- relaxed loads --
    // and a record digested from a page that has changed since is caught by the window check.
    auto digest = [&](int cell, const uint16_t* row, uint32_t (&khi)[8], uint32_t& m8,
                      uint32_t& e8, int& rank, int& n_cand) {
      const int page = cell >> 8, off = cell & 255;
      const int base = page * 256 + 8 * lane;
      const uint4 ndv = ld_rlx_v4(row + 8 * lane);
      const uint4 dwv = ld_rlx_v4(&sm.dw[base]);
      const uint4 pu0 = ld_rlx_v4(&sm.prio[base]);
      const uint4 pu1 = ld_rlx_v4(&sm.prio[base + 4]);
      const int4 pr0 = make_int4((int)pu0.x, (int)pu0.y, (int)pu0.z, (int)pu0.w);
      const int4 pr1 = make_int4((int)pu1.x, (int)pu1.y, (int)pu1.z, (int)pu1.w);
      const uint32_t nd[8] = {ndv.x & 0xffffu, ndv.x >> 16, ndv.y & 0xffffu, ndv.y >> 16,
                              ndv.z & 0xffffu, ndv.z >> 16, ndv.w & 0xffffu, ndv.w >> 16};
      const uint32_t dwl[8] = {dwv.x & 0xffffu, dwv.x >> 16, dwv.y & 0xffffu, dwv.y >> 16,
                               dwv.z & 0xffffu, dwv.z >> 16, dwv.w & 0xffffu, dwv.w >> 16};
      const int pr[8] = {pr0.x, pr0.y, pr0.z, pr0.w, pr1.x, pr1.y, pr1.z, pr1.w};
      // the popped cell itself has its diff weight and priority zeroed first
      // (video.py:140-141): never a candidate
      const uint32_t not_mine = lane == (off >> 3) ? ~(1u << (off & 7)) : ~0u;
      m8 = 0;
      e8 = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int delta = (int)nd[j] - (int)dwl[j];
        khi[j] = (uint32_t)(delta + 32768) & 0xffffu;
        m8 |= (delta < 0 ? 1u : 0u) << j;
        e8 |= (pr[j] != 0 ? 1u : 0u) << j;
      }
      m8 &= not_mine;
      e8 &= not_mine;
      rank = 0;
      n_cand = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t bal = __ballot_sync(kFull, (m8 >> j) & 1u);
        rank += __popc(bal & lt_mask);
        n_cand += __popc(bal);
      }
    };
    // First live entry of the sorted heap at or after `cursor` (32 entries per probe).
    auto probe = [&](int cursor) -> int {
      while (cursor < n_first) {
        const int idx = cursor + lane;
        bool live = false;
        if (idx < n_first)
          live = ld_rlx_u32(&sm.prio[(int)(sm.keys[idx] & 0x1fffu)]) != 0;
        const uint32_t bal = __ballot_sync(kFull, live);   // video.py:130
        if (bal) return cursor + __ffs(bal) - 1;
        cursor += 32;
      }
      return -1;
    };

    if (warp == kDecideWarp) {
      int n_pushed = 0, mt_issued = 0, mt_seen = 0;
      int r = 0;                 // next front-end record
      bool heap_done = false;    // sorted heap exhausted: re-queued cells only
      uint32_t hist = 0xffu;     // lane l < 16: page of the record r' = l (mod 16) decided last
#ifdef IIV_X_TIMING
      long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      int tcnt[4] = {0, 0, 0, 0};     // settled, contenders, redigest, dead
      long long tprev = clock64();
#define IIV_T(k) { const long long tn = clock64(); tacc[k] += tn - tprev; tprev = tn; }
#else
#define IIV_T(k)
#endif
      while (emitted < budget) {
        IIV_T(0)
        // ---- stream P bookkeeping: kPyBlocks resident 624-word blocks ------------------
        // Moving on to block c frees the slot of block c - 1, into which the twister makes
        // block c + 3 (request number = blocks left behind).  Reads run over into block
        // c + 1 at most, which was requested two moves ago: one request may be outstanding.
        // (Every lane polls the same shared word in one broadcast load, so the loop
        // conditions below are warp-uniform.)
        if (pos_py > 624) {
          if (mt_seen < mt_issued - 1) {
            const long long c0 = clock64();
            // back off rather than spin (the helper that makes the block is on another
            // scheduler, but a spin costs power and the wake-up is not on a tight path)
            // acquire: the block's nonce bytes are read after the flag
            while ((mt_seen = (int)ld_acq_u32(&sm.mt_done)) < mt_issued - 1) __nanosleep(40);
            wait_mt += clock64() - c0;
          }
          py_cur = (py_cur + 1) & (kPyBlocks - 1);
          pos_py -= 624;
          ++mt_issued;
          // release: this warp's reads of the block being replaced come first
          if (lane == 0) st_rel_u32(&sm.mt_req, (uint32_t)mt_issued);
        }
        const uint8_t* nonces = sm.py_nonce + py_cur * 624 + pos_py;

        int cell, slot, rank, n_cand, n_cont = -1;   // n_cont < 0: evaluate all 256 offsets
        uint32_t content, m8, e8, khi[8], cont = 0;
        uint32_t settled = 0, settled_p = 0;         // words z, w of a front-end record
        bool have = false, redigest = false;
        if (!heap_done) {
          // ---- next record of the front ends -------------------------------------------
          const int rs = r % kRecRing;
          const uint32_t seq = (uint32_t)(r + 1) & 0xffffu;
          uint4 rec = ld_acq_v4(sm.rec[rs]);
          if ((rec.x >> 16) != seq) {
            const long long c0 = clock64();
            do {
              rec = ld_acq_v4(sm.rec[rs]);
            } while ((rec.x >> 16) != seq);
            wait_rows += clock64() - c0;
          }
          IIV_T(1)
          const uint32_t kind = (rec.x >> 14) & 3u;
          if (kind == kKindEndOfHeap) {
            heap_done = true;
            if (lane == 0) st_rlx_u32(&sm.head, (uint32_t)n_first);
          } else {
            const int e = (int)(rec.x & 0x3fffu);
            cell = (int)(rec.y & 0x1fffu);
            content = (rec.y >> 13) & 0xffu;
            slot = e % kRing;
            // entries before e are done with: their rows were last read before the fence
            // that preceded the latest b_done store (release pattern on this word too)
            if (lane == 0) st_rlx_u32(&sm.head, (uint32_t)e);
            // valid unless a record decided after the front end's read hit this page
            const int since = (int)((rec.z >> 22) & 15u);     // records decided since then
            const bool in_window = (lane < 16) & (((r - 1 - lane) & 15) < since);
            const bool conflict =
                __ballot_sync(kFull, in_window && hist == (uint32_t)(cell >> 8)) != 0;
            const uint32_t n_cont_rec = (rec.z >> 16) & 63u;
            if (kind == kKindDead) {
              // dropped by the front end: the cell is zero
            } else if (!conflict && n_cont_rec <= 32u) {
              n_cand = (int)((rec.y >> 21) & 0x1ffu);
              n_cont = (int)n_cont_rec;
              settled = rec.z;
              settled_p = rec.w;
              // the contenders were stored before the record word was released
              if (!((settled >> 26) & 1u)) cont = sm.rec_cont[rs][lane];
              have = true;
            } else if (ld_rlx_u32(&sm.prio[cell]) != 0) {
              redigest = true;     // below, at the one call site this warp has
              have = true;
              // a settled record is published relaxed (it carries all it needs): acquire
              // the row from its producer here
              (void)ld_acq_u32(&tags[slot]);
            }
            // else: the cell was zeroed meanwhile -- it would be popped and skipped
            if (lane == ((r & 15))) hist = have ? (uint32_t)(cell >> 8) : 0xffu;
            ++r;
            if (!have) {
              __syncwarp();
              if (lane == 0) st_rel_u32(&sm.b_done, (uint32_t)r);
              continue;
            }
          }
        }
        if (heap_done) {
          // first-pass heap exhausted: arg-min over live re-queued cells, scored on demand
          uint64_t best = kDead;
          const int n_shared = min(n_pushed, kPushedCap);
          for (int k = lane; k < n_shared; k += 32) {
            const uint64_t key = sm.pushed[k];
            if (key == kDead) continue;
            if (sm.prio[key & 0x1fffu] == 0) {
              sm.pushed[k] = kDead;  // stale: would be popped and skipped
              continue;
            }
            best = key < best ? key : best;
          }
          for (int k = kPushedCap + lane; k < n_pushed; k += 32) {   // overflow list
            const uint64_t key = ovf[k - kPushedCap];
            if (key == kDead) continue;
            if (sm.prio[key & 0x1fffu] == 0) {
              ovf[k - kPushedCap] = kDead;
              continue;
            }
            best = key < best ? key : best;
          }
          best = warp_min64(best);
          if (best == kDead) {
            out_of_work = true;
            break;
          }
          for (int k = lane; k < n_shared; k += 32)
            if (sm.pushed[k] == best) sm.pushed[k] = kDead;
          for (int k = kPushedCap + lane; k < n_pushed; k += 32)
            if (ovf[k - kPushedCap] == best) ovf[k - kPushedCap] = kDead;
          cell = (int)(best & 0x1fffu);
          slot = kRing;
          content = __ldg(tmem + cell);
          score_row<MODE>(tp + (cell >> 8) * 128, table, content, is_aux, lane,
                          sm.ring_row[kRing]);
          __syncwarp();
          redigest = true;
        }
        IIV_T(2)
#ifdef IIV_X_TIMING
        if (redigest) ++tcnt[2]; else if ((settled >> 26) & 1u) ++tcnt[0]; else ++tcnt[1];
#endif
        if (redigest) digest(cell, sm.ring_row[slot], khi, m8, e8, rank, n_cand);
        IIV_T(3)
        const int page = cell >> 8, off = cell & 255;
        if (MODE == IIV_MODE_DHGR && content >= 0x80u) error_flags |= 1;  // :137

        // ---- _compute_error (video.py:275-301) -------------------------------------------
        // nonces of the (up to two) re-queue draws that follow the candidates' (:173-178)
        const uint32_t push_nonce0 = nonces[n_cand], push_nonce1 = nonces[n_cand + 1];
        // every candidate draws one getrandbits(8), in ascending offset order (:290-293);
        // only those with a live priority compete (:159); the two smallest (delta, nonce,
        // offset) win
        uint32_t b1, b2;
        uint32_t p1 = 0, p2 = 0;
        int o1 = off, o2 = off;
        bool has1, has2;
        if ((settled >> 26) & 1u) {
          // the two smallest deltas are unique: no nonce can change the winners, and the
          // front end has already looked up their new diffs
          has1 = (settled >> 27) & 1u;
          has2 = (settled >> 28) & 1u;
          if (has1) {
            o1 = (int)(settled & 255u);
            p1 = settled_p & 0xffffu;
          }
          if (has2) {
            o2 = (int)((settled >> 8) & 255u);
            p2 = settled_p >> 16;
          }
        } else {
        if (n_cont >= 0) {
          // front-end record: lane i holds contender i
          uint32_t key = 0xffffffffu;
          if (lane < n_cont)
            key = (cont & 0xffff00ffu) | ((uint32_t)nonces[(cont >> 8) & 0xffu] << 8);
          b1 = __reduce_min_sync(kFull, key);
          b2 = __reduce_min_sync(kFull, key == b1 ? 0xffffffffu : key);
        } else {
          uint32_t key[8];
          const uint32_t use8 = m8 & e8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t nonce = nonces[rank + __popc(m8 & ((1u << j) - 1u))];
            const uint32_t k = (khi[j] << 16) | (nonce << 8) | (uint32_t)(8 * lane + j);
            key[j] = ((use8 >> j) & 1u) ? k : 0xffffffffu;
          }
          // two smallest of the lane's keys (keys are distinct: the offset is in them)
          uint32_t lo[4], hi[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            lo[q] = min(key[2 * q], key[2 * q + 1]);
            hi[q] = max(key[2 * q], key[2 * q + 1]);
          }
          const uint32_t lo01 = min(lo[0], lo[1]),
                         hi01 = min(max(lo[0], lo[1]), min(hi[0], hi[1]));
          const uint32_t lo23 = min(lo[2], lo[3]),
                         hi23 = min(max(lo[2], lo[3]), min(hi[2], hi[3]));
          const uint32_t best1 = min(lo01, lo23);
          const uint32_t best2 = min(max(lo01, lo23), min(hi01, hi23));
          b1 = __reduce_min_sync(kFull, best1);
          b2 = __reduce_min_sync(kFull, best1 == b1 ? best2 : best1);
        }
        // byte_pair_difference of the accepted offsets (video.py:166) = their new diff
        has1 = b1 != 0xffffffffu;
        has2 = b2 != 0xffffffffu;
        if (has1) {
          o1 = (int)(b1 & 255u);
          p1 = sm.ring_row[slot][o1];
        }
        if (has2) {
          o2 = (int)(b2 & 255u);
          p2 = sm.ring_row[slot][o2];
        }
        }
        IIV_T(4)
        int push1 = p1 != 0, push2 = p2 != 0;
        if (n_pushed + push1 + push2 > pushed_cap) {   // the host sizes the overflow list so
          error_flags |= 4;                            // that this cannot happen
          push1 = push2 = 0;
        }
        if (lane == 0) {
          // relaxed: the front ends and producers read these words while they change
          st_rlx_u32(&sm.prio[cell], 0u);                                  // video.py:140
          st_rlx_u16(&sm.dw[cell], 0);                                     // video.py:141
          if (has1) st_rlx_u32(&sm.prio[page * 256 + o1], p1);            // video.py:170
          if (has2) st_rlx_u32(&sm.prio[page * 256 + o2], p2);
          if (push1) {   // video.py:173-178
            const uint64_t key = requeue_key(p1, push_nonce0, page * 256 + o1);
            if (n_pushed < kPushedCap) sm.pushed[n_pushed] = key;
            else ovf[n_pushed - kPushedCap] = key;
          }
          if (push2) {
            const uint64_t key =
                requeue_key(p2, push1 ? push_nonce1 : push_nonce0, page * 256 + o2);
            const int at = n_pushed + push1;
            if (at < kPushedCap) sm.pushed[at] = key;
            else ovf[at - kPushedCap] = key;
          }
          // video.py:185-187: pad to 4 with offsets[0]
          uint2 rec;
          rec.x = (uint32_t)(page + 32) | (content << 8) | ((uint32_t)off << 16) |
                  ((uint32_t)o1 << 24);
          rec.y = (uint32_t)o2 | ((uint32_t)off << 8) | (1u << 16);
          const uint2 rec_out = rec;
          *reinterpret_cast<uint2*>(seg_out + (size_t)emitted * 8) = rec_out;
          // hand the stores to the applier (video.py:144, :172); the queue is checked
          // for room once per 16 records
          if ((emitted & 15) == 0 &&
              emitted - (int)ld_acq_u32(&sm.applied_pub) > kOpQueue - 16) {
            const long long c0 = clock64();
            // the applier runs on another scheduler; back off all the same
            while (emitted - (int)ld_acq_u32(&sm.applied_pub) > kOpQueue - 16) __nanosleep(20);
            wait_misc += clock64() - c0;
          }
          rec.y |= (uint32_t)((emitted + 1) & 255) << 24;
          // one word, its own ready flag (sequence number in the top byte)
          st_rlx_u64(&sm.opq[emitted % kOpQueue], ((unsigned long long)rec.y << 32) | rec.x);
          // the priorities above are in place: whoever sees this b_done sees them, and what
          // this warp read of the record's row and contenders came first
          if (!heap_done) {
            fence_cta();
            st_rlx_u32(&sm.b_done, (uint32_t)r);
          }
        }
        IIV_T(5)
        n_pushed += push1 + push2;
        pos_py += n_cand + push1 + push2;
        py_words += n_cand + push1 + push2;
        ++emitted;
        __syncwarp();
        IIV_T(6)
      }
#ifdef IIV_X_TIMING
      if (lane == 0 && clip == 0)
        printf("seg %d phaseA: gathers %lld fold+rank %lld nonces %lld keys %lld select %lld sort %lld\n", seg, pa_sh[6], pa_sh[1], pa_sh[2], pa_sh[3], pa_sh[4], pa_sh[5]);
      if (lane == 0 && clip == 0)
        printf("seg %d emitted %d | top %lld recwait %lld classify %lld digest %lld winners %lld stores %lld sync %lld | settled %d cont %d redigest %d\n",
               seg, emitted, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6],
               tcnt[0], tcnt[1], tcnt[2]);
#endif
      while (mt_seen < mt_issued) mt_seen = (int)ld_acq_u32(&sm.mt_done);
      if (lane == 0) {
        st_rlx_u32(&sm.final_emitted, (uint32_t)emitted);
        sm.wmin64[0] = (uint64_t)wait_rows;
        sm.wmin64[1] = (uint64_t)(wait_misc + wait_mt);
        sm.wmin64[2] = (uint64_t)(clock64() - clk_b);
        sm.scan[0] = emitted;
        sm.scan[1] = out_of_work ? 1 : 0;
        sm.scan[2] = py_words;
        sm.scan[3] = pos_py;
        sm.scan[4] = py_cur;
        sm.scan[5] = error_flags;
        st_rel_u32(&sm.stop, 1u);      // final_emitted is in place for whoever sees it
      }
    } else if (warp == kDecideWarp - 1 || warp == kDecideWarp - 2) {
      // ---- front end: records r = f, f + 2, ... ----------------------------------------
      const int f = kDecideWarp - 1 - warp;
      for (int r = f;; r += kFronts) {
        // my turn to pop, with room in the record ring?
        bool quit = false;
        uint32_t pop = 0;
        while (true) {
          if (ld_rlx_u32(&sm.stop)) {
            quit = true;
            break;
          }
          pop = ld_rlx_u32(&sm.pop_state);
          // acquire: record slot r % kRecRing (word and contenders) is rewritten below, after
          // the decision warp's reads of record r - kRecRing
          if ((int)(pop >> 14) == r && r < (int)ld_acq_u32(&sm.b_done) + kRecRing) break;
          __nanosleep(20);
        }
        if (quit) break;
        const int rs = r % kRecRing;
        const uint32_t seq = ((uint32_t)(r + 1) & 0xffffu) << 16;
        const int e = probe((int)(pop & 0x3fffu));
        if (e < 0) {
          // heap exhausted: tell the decision warp, and let the other front end see it too
          if (lane == 0) {
            st_rlx_v4(sm.rec[rs], make_uint4(seq | (kKindEndOfHeap << 14), 0u, 0u, 0u));
            st_rlx_u32(&sm.pop_state, ((uint32_t)(r + 1) << 14) | (pop & 0x3fffu));
          }
          break;
        }
        const int cell = (int)(sm.keys[e] & 0x1fffu);
        // sampled BEFORE the page is read (an acquire keeps the later loads behind it): every
        // opcode that can have changed the page since is then inside the window the decision
        // warp checks
        const int seen = (int)ld_acq_u32(&sm.b_done);
        if (lane == 0) st_rlx_u32(&sm.pop_state, ((uint32_t)(r + 1) << 14) | (uint32_t)(e + 1));
        // the row of this entry, from the producers.  If the cell has been zeroed since the
        // probe (by an opcode decided meanwhile) nobody may ever score it: hand over a
        // dead record, which the decision warp drops like the pop-and-skip it stands for.
        const int slot = e % kRing;
        uint32_t tag = 0;
        bool dead = false;
        while (true) {
          if (ld_rlx_u32(&sm.prio[cell]) == 0) {
            dead = true;
            break;
          }
          // Flow control of the row ring: the producers may fill [head, head + kRing).
          // Once every earlier record has been decided this entry is the oldest one in
          // use, however many dead entries precede it -- move the window up to it, or
          // nobody would ever score it.
          if (lane == 0 && (int)ld_acq_u32(&sm.b_done) == r) st_rlx_u32(&sm.head, (uint32_t)e);
          tag = ld_acq_u32(&tags[slot]);       // acquire: the row is read after its tag
          if ((tag >> 8) == (uint32_t)e) break;
          if (ld_rlx_u32(&sm.stop)) {
            quit = true;
            break;
          }
        }
        if (quit) break;
        if (dead) {
          if (lane == 0)
            st_rlx_v4(sm.rec[rs], make_uint4(seq | (kKindDead << 14) | (uint32_t)e,
                                             (uint32_t)cell, 0u, 0u));
          continue;
        }
        // The row is read after its tag and the page after `seen`: both were acquire loads
        // (plain LDS on this part; a fence here cost 5 % of a clip's time).
        const uint16_t* row = sm.ring_row[slot];
        uint32_t khi[8], m8, e8;
        int rank, n_cand;
        digest(cell, row, khi, m8, e8, rank, n_cand);
        // Whatever the nonces, the two winners have one of the two smallest deltas among
        // the competing candidates: pass only those on (with their nonce ranks).
        const uint32_t use8 = m8 & e8;
        // The lane's two smallest (delta, offset) keys among its competing candidates, then
        // the warp's two smallest, g1 < g2.
        uint32_t mine_key = 0xffffffffu, mine_key2 = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if ((use8 >> j) & 1u) {
            const uint32_t k = (khi[j] << 16) | (uint32_t)(8 * lane + j);
            mine_key2 = min(mine_key2, max(mine_key, k));
            mine_key = min(mine_key, k);
          }
        }
        const uint32_t g1 = __reduce_min_sync(kFull, mine_key);
        const uint32_t g2 = __reduce_min_sync(kFull, mine_key == g1 ? mine_key2 : mine_key);
        // With at most one candidate at each of the two smallest deltas the nonces cannot
        // change the outcome (video.py:295-301 sorts by delta first): the winners are g1 and
        // g2, and their new diffs are looked up here, off the decision warp's chain.  That is
        // the case when g2 does not share g1's delta and nobody else shares g2's.
        // Otherwise the contenders are handed over with their nonce ranks.
        uint32_t settled = 0, settled_p = 0;
        int n_cont = 0;
        bool unique = g2 == 0xffffffffu;
        if (!unique && (g1 >> 16) != (g2 >> 16)) {
          bool rival = false;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            rival |= ((use8 >> j) & 1u) && khi[j] == (g2 >> 16) &&
                     (uint32_t)(8 * lane + j) != (g2 & 0xffffu);
          unique = __ballot_sync(kFull, rival) == 0;
        }
        if (unique) {
          settled = 1u << 26;
          if (g1 != 0xffffffffu) {
            const uint32_t w1 = g1 & 255u;
            settled |= w1 | (1u << 27);
            settled_p = row[w1];
            n_cont = 1;
          }
          if (g2 != 0xffffffffu) {
            const uint32_t w2 = g2 & 255u;
            settled |= (w2 << 8) | (1u << 28);
            settled_p |= (uint32_t)row[w2] << 16;
            n_cont = 2;
          }
        } else {
          // whatever the nonces, the two winners have one of the two smallest deltas
          const uint32_t d1 = g1 >> 16;
          uint32_t d2 = g2 >> 16;
          if (d2 == d1) {
            uint32_t lane_min2 = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (((use8 >> j) & 1u) && khi[j] > d1) lane_min2 = min(lane_min2, khi[j]);
            d2 = __reduce_min_sync(kFull, lane_min2);
          }
          uint32_t c8 = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (((use8 >> j) & 1u) && khi[j] <= d2) c8 |= 1u << j;
          const int mine = __popc(c8);
          int incl = mine;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += v;
          }
          n_cont = __shfl_sync(kFull, incl, 31);
          if (n_cont <= 32) {
            int at = incl - mine;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if ((c8 >> j) & 1u) {
                const uint32_t nrank = (uint32_t)(rank + __popc(m8 & ((1u << j) - 1u)));
                sm.rec_cont[rs][at++] = (khi[j] << 16) | (nrank << 8) | (uint32_t)(8 * lane + j);
              }
            }
          }
        }
        // the contenders (all lanes), if any, go first, then the record word: a release when
        // something sits behind it (the contenders; the row, which the decision warp then
        // reads on the strength of this warp's acquire of its tag), relaxed when the word is
        // all there is
        __syncwarp();
        if (lane == 0) {
          const uint4 word = make_uint4(
              seq | (kKindLive << 14) | (uint32_t)e,
              (uint32_t)cell | ((tag & 0xffu) << 13) | ((uint32_t)n_cand << 21),
              settled | ((uint32_t)min(n_cont, 63) << 16) | ((uint32_t)(r - seen) << 22),
              settled_p);
          if (settled) st_rlx_v4(sm.rec[rs], word);
          else st_rel_v4(sm.rec[rs], word);
        }
      }
    } else if (warp < kProducers) {
      // ---- producers: entries 2p, 2p+1 (mod 2 * kProducers), two rows in flight ----------
      for (int e0 = 2 * warp; e0 < n_first; e0 += 2 * kProducers) {
        const int e1 = e0 + 1;
        bool quit = false;
        while (true) {
          int h = 0, st = 0;
          if (lane == 0) {
            h = (int)ld_acq_u32(&sm.head);     // acquire: rows of entries before h are free
            st = (int)ld_rlx_u32(&sm.stop);
          }
          h = __shfl_sync(kFull, h, 0);
          st = __shfl_sync(kFull, st, 0);
          if (st) {
            quit = true;
            break;
          }
          if (e1 < h + kRing) break;
          __nanosleep(100);
        }
        if (quit) break;
        const int cell0 = (int)(sm.keys[e0] & 0x1fffu);
        const int cell1 = e1 < n_first ? (int)(sm.keys[e1] & 0x1fffu) : cell0;
        const bool alive0 = ld_rlx_u32(&sm.prio[cell0]) != 0;
        const bool alive1 = e1 < n_first && ld_rlx_u32(&sm.prio[cell1]) != 0;
        uint32_t c0 = 0, c1 = 0;
        if (alive0) c0 = __ldg(tmem + cell0);
        if (alive1) c1 = __ldg(tmem + cell1);
        // Rows are scored into registers first.  A producer can fall a whole ring revolution
        // behind on an entry that died after its aliveness check (nobody waits for such a
        // row, so the window moves on), and the slot may then be due for entry e + kRing --
        // which is this same warp's (kRing is a multiple of the producers' round), a later
        // iteration of this loop: the rows of a slot are written in entry order, by one warp.
        static_assert(kRing % (2 * kProducers) == 0, "a ring slot must stay with one producer");
        uint4 row0 = make_uint4(0, 0, 0, 0), row1 = row0;
        if (alive0) row0 = score_row_regs<MODE>(tp + (cell0 >> 8) * 128, table, c0, is_aux, lane);
        if (alive1) row1 = score_row_regs<MODE>(tp + (cell1 >> 8) * 128, table, c1, is_aux, lane);
        if (alive0) reinterpret_cast<uint4*>(sm.ring_row[e0 % kRing])[lane] = row0;
        if (alive1) reinterpret_cast<uint4*>(sm.ring_row[e1 % kRing])[lane] = row1;
        __syncwarp();      // every lane's row stores, then lane 0's release of the tags
        if (lane == 0 && (alive0 || alive1)) {
          fence_cta();     // one fence releases both tags (a fence followed by strong stores)
          if (alive0) st_rlx_u32(&tags[e0 % kRing], ((uint32_t)e0 << 8) | c0);
          if (alive1) st_rlx_u32(&tags[e1 % kRing], ((uint32_t)e1 << 8) | c1);
        }
        __syncwarp();
      }
    } else if (warp == kHelpWarp) {
      // helper: two background jobs in one warp.
      //  * twister: prepares the next MT19937 block of stream P when the decision warp asks
      //    for it;
      //  * applier: Bitmap.apply for (off, o1, o2) of each published record.  Stores only
      //    interact inside a page (a packed word and its two neighbours), so lane i takes
      //    the i-th pending record and records of different pages are applied at once;
      //    records that share a page go in emission order, one per round.  Re-applying an
      //    offset that repeats offsets[0] is idempotent.  Nothing in phase B reads the
      //    source, and the queue is 64 records deep: a block twist in between is no matter.
      int done = 0, applied = 0;
      while (true) {
        int req = 0, st = 0, fin = 0;
        if (lane == 0) {
          req = (int)ld_acq_u32(&sm.mt_req);    // acquire: the slot to overwrite has been read
          st = (int)ld_acq_u32(&sm.stop);       // stop is set last, with a release
          fin = (int)ld_rlx_u32(&sm.final_emitted);
        }
        req = __shfl_sync(kFull, req, 0);
        st = __shfl_sync(kFull, st, 0);
        fin = __shfl_sync(kFull, fin, 0);
        if (req > done) {
          // request n (from 1) replaces the block n behind the segment's first one by the
          // block kPyBlocks - 1 + n ahead of it, made from its predecessor
          const int dst = (py_cur + done) & (kPyBlocks - 1);
          const int src = (dst + kPyBlocks - 1) & (kPyBlocks - 1);
          warp_twist<true>(sm.mt_py[src], sm.mt_py[dst], lane, dst, sm.py_nonce);
          ++done;
          __syncwarp();     // every lane's words and nonce bytes, then lane 0's release
          if (lane == 0) st_rel_u32(&sm.mt_done, (uint32_t)done);
          continue;
        }
        const int idx = applied + lane;
        const unsigned long long raw = ld_rlx_u64(&sm.opq[idx % kOpQueue]);
        const bool ready = (uint32_t)(raw >> 56) == (uint32_t)((idx + 1) & 255) && raw != 0ull;
        // the records form a prefix: lane i can only go if lanes < i are ready as well
        const uint32_t rdy = __ballot_sync(kFull, ready);
        const int n = __ffs(~rdy) - 1;          // length of the ready prefix (0..32; -1 -> 32)
        const int count = n < 0 ? 32 : n;
        if (count > 0) {
          const bool mine = lane < count;
          const uint32_t rx = (uint32_t)raw, ry = (uint32_t)(raw >> 32);
          const int page = (int)(rx & 0xffu) - 32;
          const uint32_t content = (rx >> 8) & 0xffu;
          const int off = (int)((rx >> 16) & 0xffu), o1 = (int)(rx >> 24);
          const int o2 = (int)(ry & 0xffu);
          // lanes holding the same page, in lane (= emission) order
          const uint32_t group = __match_any_sync(kFull, mine ? page : 64 + lane);
          uint32_t pending = group;
          while (__any_sync(kFull, mine && pending != 0)) {
            if (mine && pending != 0 && __ffs(pending) - 1 == lane) {   // my turn in my group
              apply_store<MODE>(sm.src, g_mem, page, off, is_aux, content);
              if (o1 != off) apply_store<MODE>(sm.src, g_mem, page, o1, is_aux, content);
              if (o2 != off) apply_store<MODE>(sm.src, g_mem, page, o2, is_aux, content);
            }
            __syncwarp();
            // every group retires its lowest pending lane
            pending &= pending - 1;
          }
          applied += count;
          // release: the queue slots just read may be rewritten by whoever sees this
          __syncwarp();
          if (lane == 0) st_rel_u32(&sm.applied_pub, (uint32_t)applied);
          continue;
        }
        if (st && applied >= fin) break;
        __nanosleep(100);
      }
    } else if (warp == kNpWarp) {
      // the stream N blocks the next heapify will draw from (one nonce per nonzero priority,
      // video.py:259-267: 13 blocks for a full screen), one after the other, then nothing:
      // this warp shares the decision warp's scheduler and must not poll
      const int np_goal = (seg + 1 < n_segments && P <= kCells / 2) ? kNpPreMax : 0;
      for (int np_made = 0; np_made < np_goal; ++np_made) {
        int st = 0;
        if (lane == 0) st = (int)ld_rlx_u32(&sm.stop);
        if (__shfl_sync(kFull, st, 0)) break;
        warp_twist<false>(np_made == 0 ? sm.mt_np[np_cur] : np_pre_blocks + (np_made - 1) * 624,
                          np_pre_blocks + np_made * 624, lane, 0, sm.py_nonce);
        if (lane == 0) st_rlx_u32(&sm.np_pre, (uint32_t)(np_made + 1));
      }
    }
    __syncthreads();
    np_ready = sm.np_pre;
    emitted = sm.scan[0];
    out_of_work = sm.scan[1] != 0;
    py_words = sm.scan[2];
    pos_py = sm.scan[3];
    py_cur = sm.scan[4];
    error_flags |= sm.scan[5];
    __syncthreads();

    // out of work: pad forever with (32, target[0,0], [0,0,0,0]) (video.py:249-251)
    if (emitted < budget) {
      const uint32_t c0 = tmem[0];
      uint2 rec;
      rec.x = 32u | (c0 << 8);
      rec.y = 0u;
      for (int k = emitted + t; k < budget; k += kThreads)
        *reinterpret_cast<uint2*>(seg_out + (size_t)k * 8) = rec;
    }
    if (t == 0) {
      info[0] = emitted;
      info[1] = prio_total;
      info[2] = n_heap;
      info[3] = py_words;
      info[4] = (int64_t)(clk_b - clk_seg);        // cycles: score + heapify
      info[5] = (int64_t)sm.wmin64[2];             // cycles: opcode loop
      info[6] = (int64_t)sm.wmin64[0];             // consumer cycles waiting for rows
      info[7] = (int64_t)sm.wmin64[1];             // decision-warp cycles waiting for MT / applier
      if (out_of_work) g_flags[is_aux ? 1 : 0] = 1;   // video.py:189
    }
    if (extras.opcodes_mirror != nullptr) {
      // second copy of the segment's records for a caller that reads them on the host (a
      // page-locked buffer mapped into the device's address space): coalesced, all threads,
      // off the opcode loop -- the decision warp itself only ever stores to device memory
      __syncthreads();
      __threadfence_block();
      uint2* dst = reinterpret_cast<uint2*>(extras.opcodes_mirror) + (seg_out - opcodes) / 8;
      const volatile uint2* srcp = reinterpret_cast<const volatile uint2*>(seg_out);
      for (int k = t; k < budget; k += kThreads) {
        uint2 v;
        v.x = srcp[k].x;
        v.y = srcp[k].y;
        dst[k] = v;
      }
    }
    op_out += (size_t)budget * 8;
    __syncthreads();
    for (int k = t; k < kCells; k += kThreads) g_prio[k] = sm.prio[k];
    __syncthreads();
  }

  // ---- store persistent state ----------------------------------------------------
  if (pos_py > 624) {  // normalise so that (state, pos) is a legal MT19937 state
    py_cur = (py_cur + 1) & (kPyBlocks - 1);
    pos_py -= 624;
  }
  for (int k = t; k < kCols; k += kThreads) g_packed[k] = sm.src[k];
  for (int k = t; k < 624; k += kThreads) {
    g_mt_np[k] = sm.mt_np[np_cur][k];
    g_mt_py[k] = sm.mt_py[py_cur][k];
  }
  if (t == 0) {
    g_mt_np[624] = (uint32_t)pos_np;
    g_mt_py[624] = (uint32_t)pos_py;
    if (error_flags) atomicOr(&g_flags[2], error_flags);
  }
  if (extras.tail_out != nullptr) {
    // the same tail once more, for a caller that wants it without a copy of its own (a
    // page-locked host buffer mapped into the device's address space)
    __syncthreads();
    uint32_t* tail = reinterpret_cast<uint32_t*>(extras.tail_out);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(state + kOffMtNp);
    for (int k = t; k < (int)((kStateBytes - kOffMtNp) / 4); k += kThreads) tail[k] = src[k];
  }
}

__global__ void mt_draw_kernel(uint32_t* mt625, uint32_t* words, int n) {
  __shared__ uint32_t buf[2][624];
  const int t = threadIdx.x;
  for (int k = t; k < 624; k += blockDim.x) buf[0][k] = mt625[k];
  int pos = (int)mt625[624];
  int cur = 0;
  __syncthreads();
  int done = 0;
  while (done < n) {
    if (pos >= 624) {
      twist(buf[cur], buf[cur ^ 1]);
      cur ^= 1;
      pos = 0;
    }
    const int take = min(n - done, 624 - pos);
    for (int k = t; k < take; k += blockDim.x)
      words[done + k] = mt_temper(buf[cur][pos + k]);
    done += take;
    pos += take;
  }
  __syncthreads();
  for (int k = t; k < 624; k += blockDim.x) mt625[k] = buf[cur][k];
  if (t == 0) mt625[624] = (uint32_t)pos;
}

}  // namespace
}  // namespace iiv

using namespace iiv;

extern "C" size_t iiv_clip_state_bytes(void) { return kStateBytes; }

extern "C" int iiv_clip_state_layout(size_t* offsets8) {
  IIV_REQUIRE(offsets8, "null pointer");
  const size_t o[IIV_CLIP_STATE_FIELDS] = {kOffPacked,  kOffMain,  kOffAux,  kOffPrioMain,
                                           kOffPrioAux, kOffMtNp,  kOffMtPy, kOffFlags};
  for (int k = 0; k < IIV_CLIP_STATE_FIELDS; ++k) offsets8[k] = o[k];
  return 0;
}

static int check_segments(int mode, int n_frames, const int32_t* h_segments, int n_segments,
                          int64_t* total_out, int* max_budget_out) {
  int64_t total = 0;
  int max_budget = 0;
  for (int s = 0; s < n_segments; ++s) {
    const int32_t* q = h_segments + 3 * s;
    IIV_REQUIRE(q[0] >= 0 && q[0] < n_frames, "segment %d: frame %d out of range", s, q[0]);
    IIV_REQUIRE(!(q[1] && mode == IIV_MODE_HGR), "segment %d: HGR has no aux bank", s);
    IIV_REQUIRE(q[2] >= 0 && q[2] <= kMaxBudget, "segment %d: budget %d outside 0..%d", s, q[2], kMaxBudget);
    total += q[2];
    if (q[2] > max_budget) max_budget = q[2];
  }
  *total_out = total;
  *max_budget_out = max_budget;
  return 0;
}

static cudaError_t scratch_pool(cudaMemPool_t* out);

static int launch_encode(int mode, int n_clips, uint8_t* d_state, size_t state_stride,
                         const uint8_t* d_target_mem, const uint64_t* d_target_packed,
                         int n_frames, const int32_t* d_segments, int n_segments, int64_t total,
                         int max_budget, const uint16_t* d_table, uint8_t* d_opcodes,
                         int64_t* d_seg_info, cudaStream_t st,
                         EncodeExtras extras = EncodeExtras{}) {
  const size_t smem = sizeof(Smem);
  cudaError_t e;
  // An opcode re-queues at most two cells (video.py:173-178): a segment of more than
  // kPushedCap / 2 opcodes gets a per-clip overflow list, stream-ordered scratch from the
  // library's own pool.
  uint64_t* d_overflow = nullptr;
  int overflow_cap = 0;
  if (2 * (int64_t)max_budget > kPushedCap) {
    overflow_cap = 2 * max_budget - kPushedCap;
    cudaMemPool_t pool;
    IIV_CUDA(scratch_pool(&pool));
    IIV_CUDA(cudaMallocFromPoolAsync((void**)&d_overflow,
                                     sizeof(uint64_t) * (size_t)overflow_cap * (size_t)n_clips,
                                     pool, st));
  }
  if (mode == IIV_MODE_HGR) {
    e = cudaFuncSetAttribute(encode_kernel<IIV_MODE_HGR>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      encode_kernel<IIV_MODE_HGR><<<n_clips, kThreads, smem, st>>>(
          d_state, state_stride, d_target_mem, d_target_packed, n_frames, d_segments,
          n_segments, d_table, d_opcodes, total, d_seg_info, d_overflow, overflow_cap,
          extras);
  } else {
    e = cudaFuncSetAttribute(encode_kernel<IIV_MODE_DHGR>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      encode_kernel<IIV_MODE_DHGR><<<n_clips, kThreads, smem, st>>>(
          d_state, state_stride, d_target_mem, d_target_packed, n_frames, d_segments,
          n_segments, d_table, d_opcodes, total, d_seg_info, d_overflow, overflow_cap,
          extras);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (d_overflow) cudaFreeAsync(d_overflow, st);
  if (e != cudaSuccess) return cuda_fail(e, "encode_kernel");
  return 0;
}

static int check_encode_args(int mode, int n_clips, const void* d_state, size_t state_stride,
                             const void* d_target_mem, const void* d_target_packed,
                             int n_frames, int n_segments, const void* d_table,
                             const void* d_seg_info) {
  IIV_REQUIRE(mode == IIV_MODE_HGR || mode == IIV_MODE_DHGR, "bad mode %d", mode);
  IIV_REQUIRE(n_clips >= 0 && n_frames > 0 && n_segments >= 0, "bad counts");
  IIV_REQUIRE(d_state && d_target_mem && d_target_packed && d_table && d_seg_info,
              "null pointer");
  IIV_REQUIRE(state_stride >= kStateBytes && state_stride % 16 == 0,
              "state_stride %zu too small or unaligned", state_stride);
  return 0;
}

// Stream-ordered scratch for the schedule of an unplanned call.  The device's default
// pool hands its pages back to the OS at every synchronisation (release threshold 0), which
// turns the next cudaMallocAsync into a real allocation of a millisecond or more; a pool of
// our own that keeps what it has makes the call allocation-free after the first one.
static cudaError_t scratch_pool(cudaMemPool_t* out) {
  constexpr int kMaxDevices = 64;
  static std::mutex mu;
  static cudaMemPool_t pools[kMaxDevices] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    e = cudaMemPoolCreate(&pools[dev], &props);
    if (e != cudaSuccess) return e;
    uint64_t keep = ~0ull;
    e = cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    if (e != cudaSuccess) return e;
  }
  *out = pools[dev];
  return cudaSuccess;
}

extern "C" int iiv_encode_clips(int mode, int n_clips, uint8_t* d_state,
                                size_t state_stride, const uint8_t* d_target_mem,
                                const uint64_t* d_target_packed, int n_frames,
                                const int32_t* h_segments, int n_segments,
                                const uint16_t* d_table, uint8_t* d_opcodes,
                                int64_t* d_seg_info, void* stream) {
  int rc = check_encode_args(mode, n_clips, d_state, state_stride, d_target_mem,
                             d_target_packed, n_frames, n_segments, d_table, d_seg_info);
  if (rc) return rc;
  IIV_REQUIRE(h_segments || n_segments == 0, "null pointer");
  if (n_clips == 0 || n_segments == 0) return 0;
  int64_t total = 0;
  int max_budget = 0;
  rc = check_segments(mode, n_frames, h_segments, n_segments, &total, &max_budget);
  if (rc) return rc;
  IIV_REQUIRE(d_opcodes || total == 0, "null opcode buffer");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* d_segments = nullptr;
  cudaMemPool_t pool;
  IIV_CUDA(scratch_pool(&pool));
  IIV_CUDA(cudaMallocFromPoolAsync((void**)&d_segments, sizeof(int32_t) * 3 * n_segments, pool,
                                   st));
  cudaError_t e = cudaMemcpyAsync(d_segments, h_segments, sizeof(int32_t) * 3 * n_segments,
                                  cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    rc = launch_encode(mode, n_clips, d_state, state_stride, d_target_mem, d_target_packed,
                       n_frames, d_segments, n_segments, total, max_budget, d_table, d_opcodes,
                       d_seg_info, st);
  cudaFreeAsync(d_segments, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(segments)");
  return rc;
}

extern "C" int iiv_encode_clips_planned(int mode, int n_clips, uint8_t* d_state,
                                        size_t state_stride, const uint8_t* d_target_mem,
                                        const uint64_t* d_target_packed, int n_frames,
                                        const int32_t* h_segments, const int32_t* d_segments,
                                        int n_segments, const uint16_t* d_table,
                                        uint8_t* d_opcodes, int64_t* d_seg_info, void* stream) {
  int rc = check_encode_args(mode, n_clips, d_state, state_stride, d_target_mem,
                             d_target_packed, n_frames, n_segments, d_table, d_seg_info);
  if (rc) return rc;
  IIV_REQUIRE((h_segments && d_segments) || n_segments == 0, "null pointer");
  if (n_clips == 0 || n_segments == 0) return 0;
  int64_t total = 0;
  int max_budget = 0;
  rc = check_segments(mode, n_frames, h_segments, n_segments, &total, &max_budget);
  if (rc) return rc;
  IIV_REQUIRE(d_opcodes || total == 0, "null opcode buffer");
  return launch_encode(mode, n_clips, d_state, state_stride, d_target_mem, d_target_packed,
                       n_frames, d_segments, n_segments, total, max_budget, d_table, d_opcodes,
                       d_seg_info, (cudaStream_t)stream);
}

// One encode_frame generator of the Python facade as ONE launch (the facade's per-generator
// cost is host time and launch gaps: a dozen separate tensor operations and copies cost more
// than the kernel's own fixed costs).  The kernel starts from d_state_in, leaves the state in
// d_state_out, and writes its results straight into the caller's page-locked host buffers
// through their device mappings: no copy is enqueued at all.
extern "C" int iiv_encode_generator(int mode, const uint8_t* d_state_in, uint8_t* d_state_out,
                                    const uint8_t* d_target_mem,
                                    const uint64_t* d_target_packed, int is_aux, int budget,
                                    const uint16_t* d_table, uint8_t* d_opcodes,
                                    uint8_t* h_opcodes, int64_t* h_seg_info,
                                    uint8_t* h_state_tail, void* event, void* stream) {
  IIV_REQUIRE(d_opcodes && h_opcodes && h_seg_info && h_state_tail && d_state_in,
              "null pointer");
  IIV_REQUIRE(!(is_aux && mode == IIV_MODE_HGR), "HGR has no aux bank");
  IIV_REQUIRE(budget >= 1 && budget <= kMaxBudget, "budget %d outside 1..%d", budget, kMaxBudget);
  // device views of the host buffers (identical under unified addressing, but ask)
  uint8_t* m_opcodes = nullptr;
  int64_t* m_info = nullptr;
  uint8_t* m_tail = nullptr;
  IIV_CUDA(cudaHostGetDevicePointer((void**)&m_opcodes, h_opcodes, 0));
  IIV_CUDA(cudaHostGetDevicePointer((void**)&m_info, h_seg_info, 0));
  IIV_CUDA(cudaHostGetDevicePointer((void**)&m_tail, h_state_tail, 0));
  int rc = check_encode_args(mode, 1, d_state_out, kStateBytes, d_target_mem, d_target_packed,
                             1, 1, d_table, m_info);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  EncodeExtras extras = {};
  extras.inline_segment = make_int3(0, is_aux ? 1 : 0, budget);
  extras.state_in = d_state_in;
  extras.tail_out = m_tail;
  extras.opcodes_mirror = m_opcodes;
  rc = launch_encode(mode, 1, d_state_out, kStateBytes, d_target_mem, d_target_packed, 1,
                     nullptr, 1, budget, budget, d_table, d_opcodes, m_info, st, extras);
  if (rc) return rc;
  if (event) IIV_CUDA(cudaEventRecord((cudaEvent_t)event, st));
  return 0;
}

extern "C" void* iiv_event_create(void) {
  cudaEvent_t ev = nullptr;
  if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  return ev;
}

extern "C" int iiv_event_wait(void* event) {
  IIV_REQUIRE(event, "null event");
  IIV_CUDA(cudaEventSynchronize((cudaEvent_t)event));
  return 0;
}

extern "C" int iiv_event_destroy(void* event) {
  if (event) IIV_CUDA(cudaEventDestroy((cudaEvent_t)event));
  return 0;
}

extern "C" int iiv_mt_draw(uint32_t* d_mt625, uint32_t* d_words, int n, void* stream) {
  IIV_REQUIRE(d_mt625 && (d_words || n == 0) && n >= 0, "bad argument");
  mt_draw_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_mt625, d_words, n);
  IIV_LAUNCH_CHECK("mt_draw_kernel");
  return 0;
}

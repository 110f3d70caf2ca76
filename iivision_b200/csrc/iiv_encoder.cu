// Path 2b: the greedy prioritised delta coder of transcoder/video.py
// (Video.encode_frame :72-93, _index_changes :95-251, _heapify_priorities
//  :253-271, _compute_error :275-301) -- one thread block per clip.
//
// The opcode loop is a sequential dependency chain (each opcode mutates the
// source bitmap and the priorities the next pop reads), so one clip cannot use
// more than one block; the chip is filled by batching independent clips.  Inside
// a block the work per opcode is data parallel over the 256 offsets of a page:
// one table gather per thread, ballot/popc ranks for the RNG draws and
// warp-REDUX min reductions for candidate selection.
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
#include "iiv_common.cuh"

namespace iiv {
namespace {

constexpr int kThreads = 256;
constexpr int kCells = 32 * 256;       // one bank: 32 pages x 256 offsets
constexpr int kCols = 32 * 128;
constexpr int kPushedCap = 4096;       // >= 2 * max budget per segment
constexpr int kMaxBudget = kPushedCap / 2;
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
  uint32_t mt_py[2][624];      // stream P: current block and its successor
  uint32_t wcnt[8];
  uint32_t wtop[8][2];
  uint64_t wmin64[8];
  int32_t scan[kThreads / 32];
  // broadcast slots
  int sel_page, sel_off, sel_state;  // sel_state: 0 = cell, 1 = need pushed, 2 = done
  uint32_t acc_p[2];
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

template <int MODE>
__device__ __forceinline__ void apply_store(uint64_t* src, uint8_t* mem, int page,
                                            int offset, int is_aux, uint32_t value) {
  const int o = byte_offset<MODE>(offset, is_aux);
  const int c = offset >> 1;
  uint64_t* row = src + page * 128;
  const uint64_t w = masked_update<MODE>(o, row[c], value);
  row[c] = w;
  if (o == 0 && c > 0)
    row[c - 1] = (row[c - 1] & keep_low_mask<MODE>()) ^ footer_of<MODE>(w);
  else if (o == Mode<MODE>::kOffsets - 1 && c < 127)
    row[c + 1] = (row[c + 1] & keep_high_mask<MODE>()) ^ header_of<MODE>(w);
  mem[page * 256 + offset] = (uint8_t)value;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
encode_kernel(uint8_t* __restrict__ states, size_t state_stride,
              const uint8_t* __restrict__ target_mem,
              const uint64_t* __restrict__ target_packed, int n_frames,
              const int32_t* __restrict__ segments, int n_segments,
              const uint16_t* __restrict__ table, uint8_t* __restrict__ opcodes,
              int64_t total_budget, int64_t* __restrict__ seg_info) {
  using M = Mode<MODE>;
  constexpr int kBanks = MODE == IIV_MODE_DHGR ? 2 : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int clip = blockIdx.x;
  uint8_t* state = states + (size_t)clip * state_stride;
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
  int pos_np = (int)g_mt_np[624]; // 0..624
  int pos_py = (int)g_mt_py[624];
  int py_cur = 0;                 // mt_py[py_cur] current block, [py_cur^1] next
  __syncthreads();
  twist(sm.mt_py[py_cur], sm.mt_py[py_cur ^ 1]);
  int error_flags = 0;

  uint8_t* op_out = opcodes + (size_t)clip * total_budget * 8;

  for (int seg = 0; seg < n_segments; ++seg) {
    const int frame = segments[3 * seg + 0];
    const int is_aux = segments[3 * seg + 1];
    const int budget = segments[3 * seg + 2];
    int64_t* info = seg_info + ((size_t)clip * n_segments + seg) * 4;
    if (budget <= 0) {  // generator created but never pulled: no side effects
      if (t == 0) info[0] = info[1] = info[2] = info[3] = 0;
      continue;
    }
    const int bank = (MODE == IIV_MODE_DHGR && is_aux) ? 1 : 0;
    const uint64_t* tp = target_packed + ((size_t)clip * n_frames + frame) * kCols;
    const uint8_t* tmem =
        target_mem + (((size_t)clip * n_frames + frame) * kBanks + bank) * kCells;
    int32_t* g_prio = reinterpret_cast<int32_t*>(
        state + (is_aux ? kOffPrioAux : kOffPrioMain));
    uint8_t* g_mem = state + (is_aux ? kOffAux : kOffMain);

    // ======================= phase A: score + heapify ===========================
    // thread t owns cells [32t, 32t+32): page t>>3, offsets (t&7)*32 .. +31.
    const int cell0 = t * 32;
    const int page_a = t >> 3;
    const int colbase = page_a * 128 + (t & 7) * 16;
    uint32_t dwv[32];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint64_t s = sm.src[colbase + c];
      const uint64_t g = __ldg(tp + colbase + c);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int o = byte_offset<MODE>(half, is_aux);
        const uint32_t x = mask_shift<MODE>(s, o), y = mask_shift<MODE>(g, o);
        const int offset = (t & 7) * 32 + 2 * c + half;
        uint32_t d = 0;
        if (!is_hole(offset))  // video.py:111
          d = __ldg(table + ((size_t)o << (2 * M::kBits)) + ((size_t)x << M::kBits) + y);
        dwv[2 * c + half] = d;
      }
    }
    int64_t prio_sum = 0;
    int nz = 0;
    uint32_t nzmask = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      int32_t p = g_prio[cell0 + k];
      prio_sum += p;
      if (dwv[k] == 0) p = 0;          // video.py:115
      p += (int32_t)dwv[k];            // video.py:116
      sm.prio[cell0 + k] = p;
      sm.dw[cell0 + k] = (uint16_t)dwv[k];
      if (p != 0) {
        ++nz;
        nzmask |= 1u << k;
      }
    }
    // block-wide sums: exclusive scan of nz, total of prio_sum.
    int incl = nz;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
      prio_sum += __shfl_xor_sync(0xffffffffu, prio_sum, d);
    if (lane == 31) sm.scan[warp] = incl;
    if (lane == 0) sm.wmin64[warp] = (uint64_t)prio_sum;
    __syncthreads();
    int rank0 = incl - nz;
    int n_heap = 0;
    int64_t prio_total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      if (w < warp) rank0 += sm.scan[w];
      n_heap += sm.scan[w];
      prio_total += (int64_t)sm.wmin64[w];
    }
    // stream N draws: word index g = pos_np + rank; block g / 624.
    const int twists = n_heap > 0 ? (pos_np + n_heap - 1) / 624 : 0;
    for (int b = 0; b <= twists; ++b) {
      if (b > 0) {
        twist(sm.mt_np[np_cur], sm.mt_np[np_cur ^ 1]);
        np_cur ^= 1;
      }
      const int lo = pos_np + rank0, hi = lo + nz;  // my word range [lo, hi)
      if (nz > 0 && lo < (b + 1) * 624 && hi > b * 624) {
        int r = 0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          if (nzmask & (1u << k)) {
            const int g = lo + r;
            if (g / 624 == b) {
              const uint32_t nonce = mt_temper(sm.mt_np[np_cur][g - b * 624]) & 0xffu;
              sm.keys[rank0 + r] = first_pass_key(sm.prio[cell0 + k], nonce, cell0 + k);
            }
            ++r;
          }
        }
      }
    }
    if (n_heap > 0) pos_np = pos_np + n_heap - 624 * twists;
    // pad to a power of two and sort ascending.
    int P = 256;
    while (P < n_heap) P <<= 1;
    for (int k = n_heap + t; k < P; k += kThreads) sm.keys[k] = kDead;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
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
    }

    // ======================= phase B: emit opcodes ==============================
    int cursor = 0;      // thread 0 only
    int n_pushed = 0;    // uniform
    int emitted = 0;     // uniform
    int py_words = 0;    // uniform
    bool out_of_work = false;
    uint8_t* seg_out = op_out;

    while (emitted < budget) {
      // keep >= 258 words of stream P addressable from pos_py (two 624-word
      // blocks are resident); advancing lazily keeps the final (state, pos) in
      // CPython's own canonical form (pos in 1..624 once a block has been used)
      if (pos_py > 2 * 624 - 258) {
        py_cur ^= 1;
        pos_py -= 624;
        twist(sm.mt_py[py_cur], sm.mt_py[py_cur ^ 1]);
      }
      if (t == 0) {
        int state_sel = 1;
        while (cursor < n_heap) {
          const int cell = (int)(sm.keys[cursor++] & 0x1fffu);
          if (sm.prio[cell] != 0) {   // video.py:130
            sm.sel_page = cell >> 8;
            sm.sel_off = cell & 255;
            state_sel = 0;
            break;
          }
        }
        sm.sel_state = state_sel;
      }
      __syncthreads();
      if (sm.sel_state == 1) {
        // first-pass heap exhausted: arg-min over live re-queued cells.
        uint64_t best = kDead;
        for (int k = t; k < n_pushed; k += kThreads) {
          const uint64_t key = sm.pushed[k];
          if (key == kDead) continue;
          if (sm.prio[key & 0x1fffu] == 0) {
            sm.pushed[k] = kDead;  // stale: would be popped and skipped
            continue;
          }
          best = key < best ? key : best;
        }
        best = warp_min64(best);
        if (lane == 0) sm.wmin64[warp] = best;
        __syncthreads();
        best = kDead;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w)
          best = sm.wmin64[w] < best ? sm.wmin64[w] : best;
        if (best == kDead) {
          out_of_work = true;
          break;  // uniform
        }
        for (int k = t; k < n_pushed; k += kThreads)
          if (sm.pushed[k] == best) sm.pushed[k] = kDead;
        if (t == 0) {
          sm.sel_page = (int)((best & 0x1fffu) >> 8);
          sm.sel_off = (int)(best & 255u);
        }
        __syncthreads();
      }
      const int page = sm.sel_page, off = sm.sel_off;
      const uint32_t content = tmem[page * 256 + off];   // video.py:134
      if (MODE == IIV_MODE_DHGR && content >= 0x80u) error_flags |= 1;  // :137

      // ---- _compute_error: one candidate offset per thread ----------------------
      const int cell = page * 256 + t;
      uint32_t dwt = sm.dw[cell];
      if (t == off) {
        sm.prio[cell] = 0;   // video.py:140
        sm.dw[cell] = 0;     // video.py:141
        dwt = 0;
      }
      uint32_t nd = 0;
      bool cand = false;
      if (!is_hole(t)) {
        const int o = byte_offset<MODE>(t, is_aux);
        const uint64_t w = __ldg(tp + page * 128 + (t >> 1));
        const uint32_t x = mask_shift<MODE>(masked_update<MODE>(o, w, content), o);
        const uint32_t y = mask_shift<MODE>(w, o);
        nd = __ldg(table + ((size_t)o << (2 * M::kBits)) + ((size_t)x << M::kBits) + y);
        cand = (int)nd - (int)dwt < 0;   // video.py:283
      }
      const uint32_t ballot = __ballot_sync(0xffffffffu, cand);
      if (lane == 0) sm.wcnt[warp] = __popc(ballot);
      __syncthreads();
      int rank = __popc(ballot & ((1u << lane) - 1u));
      int n_cand = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) {
        if (w < warp) rank += sm.wcnt[w];
        n_cand += sm.wcnt[w];
      }
      uint32_t key = 0xffffffffu;
      if (cand && sm.prio[cell] != 0) {   // video.py:159
        const int g = pos_py + rank;
        const uint32_t word = g < 624 ? sm.mt_py[py_cur][g] : sm.mt_py[py_cur ^ 1][g - 624];
        const uint32_t nonce = mt_temper(word) >> 24;
        const int delta = (int)nd - (int)dwt;
        key = ((uint32_t)(delta + 32768) << 16) | (nonce << 8) | (uint32_t)t;
      }
      const uint32_t m1 = __reduce_min_sync(0xffffffffu, key);
      const uint32_t m2 = __reduce_min_sync(0xffffffffu, key == m1 ? 0xffffffffu : key);
      if (lane == 0) {
        sm.wtop[warp][0] = m1;
        sm.wtop[warp][1] = m2;
      }
      __syncthreads();
      uint32_t b1 = 0xffffffffu, b2 = 0xffffffffu;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const uint32_t v = sm.wtop[w][e];
          if (v < b1) {
            b2 = b1;
            b1 = v;
          } else if (v < b2) {
            b2 = v;
          }
        }
      }
      if (key != 0xffffffffu && (key == b1 || key == b2)) {
        sm.acc_p[key == b1 ? 0 : 1] = nd;   // byte_pair_difference (video.py:166)
        sm.prio[cell] = (int32_t)nd;        // video.py:170
      }
      __syncthreads();
      int pushes = 0;
      {
        // uniform bookkeeping (every thread computes the same values)
        const uint32_t p1 = b1 != 0xffffffffu ? sm.acc_p[0] : 0u;
        const uint32_t p2 = b2 != 0xffffffffu ? sm.acc_p[1] : 0u;
        const int push1 = p1 != 0, push2 = p2 != 0;
        pushes = push1 + push2;
        if (t == 0) {
          apply_store<MODE>(sm.src, g_mem, page, off, is_aux, content);  // :144
          int o1 = off, o2 = off;
          int g = pos_py + n_cand;
          if (b1 != 0xffffffffu) {
            o1 = (int)(b1 & 255u);
            apply_store<MODE>(sm.src, g_mem, page, o1, is_aux, content);  // :172
            if (push1) {
              const uint32_t word = g < 624 ? sm.mt_py[py_cur][g] : sm.mt_py[py_cur ^ 1][g - 624];
              sm.pushed[n_pushed] = requeue_key(p1, mt_temper(word) >> 24, page * 256 + o1);
              ++g;
            }
          }
          if (b2 != 0xffffffffu) {
            o2 = (int)(b2 & 255u);
            apply_store<MODE>(sm.src, g_mem, page, o2, is_aux, content);
            if (push2) {
              const uint32_t word = g < 624 ? sm.mt_py[py_cur][g] : sm.mt_py[py_cur ^ 1][g - 624];
              sm.pushed[n_pushed + push1] = requeue_key(p2, mt_temper(word) >> 24, page * 256 + o2);
            }
          }
          // video.py:185-187: pad to 4 with offsets[0]
          uint2 rec;
          rec.x = (uint32_t)(page + 32) | (content << 8) | ((uint32_t)off << 16) |
                  ((uint32_t)o1 << 24);
          rec.y = (uint32_t)o2 | ((uint32_t)off << 8) | (1u << 16);
          *reinterpret_cast<uint2*>(seg_out + (size_t)emitted * 8) = rec;
        }
        n_pushed += pushes;
        pos_py += n_cand + pushes;
        py_words += n_cand + pushes;
        ++emitted;
      }
      __syncthreads();
    }

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
      if (out_of_work) g_flags[is_aux ? 1 : 0] = 1;   // video.py:189
    }
    op_out += (size_t)budget * 8;
    __syncthreads();
    for (int k = t; k < kCells; k += kThreads) g_prio[k] = sm.prio[k];
    __syncthreads();
  }

  // ---- store persistent state ----------------------------------------------------
  if (pos_py > 624) {  // normalise so that (state, pos) is a legal MT19937 state
    py_cur ^= 1;
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

extern "C" int iiv_encode_clips(int mode, int n_clips, uint8_t* d_state,
                                size_t state_stride, const uint8_t* d_target_mem,
                                const uint64_t* d_target_packed, int n_frames,
                                const int32_t* h_segments, int n_segments,
                                const uint16_t* d_table, uint8_t* d_opcodes,
                                int64_t* d_seg_info, void* stream) {
  IIV_REQUIRE(mode == IIV_MODE_HGR || mode == IIV_MODE_DHGR, "bad mode %d", mode);
  IIV_REQUIRE(n_clips >= 0 && n_frames > 0 && n_segments >= 0, "bad counts");
  IIV_REQUIRE(d_state && d_target_mem && d_target_packed && h_segments && d_table &&
                  d_opcodes && d_seg_info, "null pointer");
  IIV_REQUIRE(state_stride >= kStateBytes && state_stride % 16 == 0,
              "state_stride %zu too small or unaligned", state_stride);
  if (n_clips == 0 || n_segments == 0) return 0;
  int64_t total = 0;
  for (int s = 0; s < n_segments; ++s) {
    const int32_t* q = h_segments + 3 * s;
    IIV_REQUIRE(q[0] >= 0 && q[0] < n_frames, "segment %d: frame %d out of range", s, q[0]);
    IIV_REQUIRE(!(q[1] && mode == IIV_MODE_HGR), "segment %d: HGR has no aux bank", s);
    IIV_REQUIRE(q[2] >= 0 && q[2] <= kMaxBudget, "segment %d: budget %d outside 0..%d", s, q[2], kMaxBudget);
    total += q[2];
  }
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* d_segments = nullptr;
  IIV_CUDA(cudaMallocAsync(&d_segments, sizeof(int32_t) * 3 * n_segments, st));
  cudaError_t e = cudaMemcpyAsync(d_segments, h_segments, sizeof(int32_t) * 3 * n_segments,
                                  cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    const size_t smem = sizeof(Smem);
    if (mode == IIV_MODE_HGR) {
      e = cudaFuncSetAttribute(encode_kernel<IIV_MODE_HGR>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess)
        encode_kernel<IIV_MODE_HGR><<<n_clips, kThreads, smem, st>>>(
            d_state, state_stride, d_target_mem, d_target_packed, n_frames, d_segments,
            n_segments, d_table, d_opcodes, total, d_seg_info);
    } else {
      e = cudaFuncSetAttribute(encode_kernel<IIV_MODE_DHGR>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess)
        encode_kernel<IIV_MODE_DHGR><<<n_clips, kThreads, smem, st>>>(
            d_state, state_stride, d_target_mem, d_target_packed, n_frames, d_segments,
            n_segments, d_table, d_opcodes, total, d_seg_info);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  cudaFreeAsync(d_segments, st);
  if (e != cudaSuccess) return cuda_fail(e, "encode_kernel");
  return 0;
}

extern "C" int iiv_mt_draw(uint32_t* d_mt625, uint32_t* d_words, int n, void* stream) {
  IIV_REQUIRE(d_mt625 && (d_words || n == 0) && n >= 0, "bad argument");
  mt_draw_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_mt625, d_words, n);
  IIV_LAUNCH_CHECK("mt_draw_kernel");
  return 0;
}

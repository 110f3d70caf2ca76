// "Next" row N2: the player byte stream of transcoder/movie.py Movie.emit_stream
// (:122-161) + opcodes.py emit_command / emit_data (:49-52, :79-89, :116-121, :144-146)
// + machine.py Machine.emit (:11-25), for a whole array of tick opcodes at once.
//
// The layout is regular: a 7-byte Header (six 0xff + video mode, no address), then 7
// bytes per tick opcode (address hi, lo of op_tick_<tick>_page_<page>, content, four
// offsets); whenever stream_pos % 2048 reaches 2044 a 4-byte Ack (address hi, lo,
// 0x54 | aux bank, 0xff) closes the 2 KiB TCP frame -- after 291 ticks in the first frame
// (the header took 7 bytes), after 292 in every later one -- and in DHGR each Ack first
// flips the MAIN/AUX bank.  Terminate (address only) and zero padding to the next 2 KiB
// boundary end the stream (a full 2 KiB of zeros when already on a boundary, :160).
// Byte and integer work, one thread per opcode, HBM-bound at 7 B out per 9 B in.
#include "iiv_common.cuh"

namespace iiv {
namespace {

constexpr int kFrame = 2048, kTick = 7, kFirst = 291, kPer = 292;

__host__ __device__ __forceinline__ size_t tick_position(size_t t) {
  // stream position of the first byte of tick opcode t
  if (t < kFirst) return kTick + kTick * t;
  const size_t u = t - kFirst;
  return (1 + u / kPer) * (size_t)kFrame + kTick * (u % kPer);
}

__global__ void emit_stream_kernel(int dhgr, const uint8_t* __restrict__ opcodes,
                                   const uint8_t* __restrict__ ticks, size_t n_ticks,
                                   const uint16_t* __restrict__ tick_addr, uint32_t ack_addr,
                                   uint32_t terminate_addr, int with_header, int video_mode,
                                   uint8_t* __restrict__ out, size_t total_len,
                                   int* __restrict__ bad) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0 && with_header) {
    for (int k = 0; k < 6; ++k) out[k] = 0xff;
    out[6] = (uint8_t)video_mode;
  }
  if (t < n_ticks) {
    const uint2 op = reinterpret_cast<const uint2*>(opcodes)[t];
    const uint32_t page = op.x & 0xffu, tick = ticks[t];
    // TICK_OPCODES keys: tick in 4..66 step 2, page in 32..63 (opcodes.py:149-165)
    if (tick < 4 || tick > 66 || (tick & 1) || page < 32 || page > 63) {
      atomicExch(bad, 1);
      return;
    }
    const uint32_t addr = tick_addr[((tick - 4) >> 1) * 32 + (page - 32)];
    const size_t pos = tick_position(t);
    uint8_t* p = out + pos;
    p[0] = (uint8_t)(addr >> 8);
    p[1] = (uint8_t)addr;
    p[2] = (uint8_t)(op.x >> 8);    // content
    p[3] = (uint8_t)(op.x >> 16);   // offsets[0..3]
    p[4] = (uint8_t)(op.x >> 24);
    p[5] = (uint8_t)op.y;
    p[6] = (uint8_t)(op.y >> 8);
    if ((pos + kTick) % kFrame == kFrame - 4) {
      // the k-th Ack (k from 1) follows k flips of the bank, starting from MAIN
      const size_t k = (pos + kTick) / kFrame + 1;
      p[7] = (uint8_t)(ack_addr >> 8);
      p[8] = (uint8_t)ack_addr;
      p[9] = (dhgr && (k & 1)) ? 0x55 : 0x54;
      p[10] = 0xff;
    }
  }
  // Terminate + zero padding
  size_t end = with_header ? (size_t)kTick : 0;
  if (n_ticks > 0) {
    end = tick_position(n_ticks - 1) + kTick;
    if (end % kFrame == kFrame - 4) end += 4;
  }
  if (t == 0) {
    out[end] = (uint8_t)(terminate_addr >> 8);
    out[end + 1] = (uint8_t)terminate_addr;
  }
  for (size_t z = end + 2 + t; z < total_len; z += (size_t)gridDim.x * blockDim.x) out[z] = 0;
}

}  // namespace
}  // namespace iiv

using namespace iiv;

extern "C" size_t iiv_stream_length(size_t n_ticks, int with_header) {
  size_t end = with_header ? (size_t)kTick : 0;
  if (n_ticks > 0) {
    end = tick_position(n_ticks - 1) + kTick;
    if (end % kFrame == kFrame - 4) end += 4;
  }
  end += 2;                                   // Terminate
  return end + (kFrame - end % kFrame);       // movie.py:160 pads a whole frame on a boundary
}

extern "C" size_t iiv_stream_ticks_within(size_t n_ticks, size_t max_bytes_out) {
  if (max_bytes_out == 0) return n_ticks;     // movie.py:133: falsy = no limit
  // ops are emitted while stream_pos < max_bytes_out at the time they are pulled; the
  // position seen by tick t includes the Ack that followed tick t-1
  size_t lo = 0, hi = n_ticks;                // first t whose position is >= max
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (tick_position(mid) >= max_bytes_out) hi = mid; else lo = mid + 1;
  }
  return lo;
}

extern "C" int iiv_emit_stream(int mode, const uint8_t* d_opcodes, const uint8_t* d_ticks,
                               size_t n_ticks, const uint16_t* d_tick_addr,
                               uint32_t ack_addr, uint32_t terminate_addr, uint8_t* d_out,
                               size_t out_capacity, int* d_bad, void* stream) {
  IIV_REQUIRE(mode == IIV_MODE_HGR || mode == IIV_MODE_DHGR, "bad mode %d", mode);
  IIV_REQUIRE(d_tick_addr && d_out && d_bad && (n_ticks == 0 || (d_opcodes && d_ticks)),
              "null pointer");
  IIV_REQUIRE(ack_addr < 65536 && terminate_addr < 65536, "opcode address out of range");
  const size_t total = iiv_stream_length(n_ticks, 1);
  IIV_REQUIRE(out_capacity >= total, "output buffer too small: %zu < %zu", out_capacity, total);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t threads = n_ticks > 2048 ? n_ticks : 2048;
  emit_stream_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
      mode == IIV_MODE_DHGR, d_opcodes, d_ticks, n_ticks, d_tick_addr, ack_addr,
      terminate_addr, 1, mode, d_out, total, d_bad);
  IIV_LAUNCH_CHECK("emit_stream_kernel");
  return 0;
}

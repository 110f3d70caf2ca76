/*
 * iivision_b200.h -- C ABI of libiivision_b200.so (sm_100a).
 *
 * The reference (KrisKennaway/ii-vision) is pure Python with no FFI; the
 * drop-in boundary is therefore the set of Python functions/methods on its two
 * hot paths.  Each entry point below is what a ctypes binding for that
 * function would call; the reference symbol it replaces is cited as
 * file:line relative to the reference tree (transcoder/...).  INTEGRATION.md
 * shows the reference-side ctypes stubs.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t, or a
 *     negative IIV_E_* code; iiv_last_error() gives a thread-local message.
 *   - "d_" pointers are device pointers, "h_" pointers are host pointers.  The
 *     caller owns every buffer; the library keeps no pointer after a call.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Work
 *     is stream-ordered; functions taking only d_ pointers do not synchronise.
 *   - mode: IIV_MODE_HGR (screen.py:550 HGRBitmap) or IIV_MODE_DHGR
 *     (screen.py:819 DHGRBitmap).
 *   - tables are uint16[n_offsets][4^masked_bits], entry index
 *     (i << masked_bits) + j  (make_data_tables.py:133-135, 163).
 */
#ifndef IIVISION_B200_H_
#define IIVISION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IIV_MODE_HGR 0
#define IIV_MODE_DHGR 1

#define IIV_E_BADARG (-1)
#define IIV_E_UNSUPPORTED (-2)
#define IIV_E_OVERFLOW (-3)

/* Table layouts for iiv_table_generate. */
#define IIV_LAYOUT_TRIANGULAR 0 /* j < i only, zeros elsewhere: the reference's
                                   .npz layout (make_data_tables.py:156-172) */
#define IIV_LAYOUT_SYMMETRIC 1  /* full square: what Bitmap.edit_distances
                                   returns after its transpose-add
                                   (screen.py:358-365) */

/* Generator kernels. */
#define IIV_ALGO_AUTO 0
#define IIV_ALGO_CHAIN 1 /* one independent 1-D recurrence per entry */
#define IIV_ALGO_TREE 2  /* shared-suffix tree over a block of j per thread */
#define IIV_ALGO_SPLIT 3 /* chain cut in two tabulated halves (what AUTO runs) */

const char* iiv_last_error(void);
int iiv_version(void);

/* Device-wide hint (cudaLimitMaxL2FetchGranularity): how much L2 fetches from HBM per miss,
 * 32, 64 or 128 bytes.  The scorer's table reads (screen.py:441, :486) are 2-byte gathers
 * spread over the whole table; 32 keeps a miss at one sector. */
int iiv_set_l2_fetch_granularity(size_t bytes);
size_t iiv_get_l2_fetch_granularity(void);

/* Measurement aid (no reference counterpart): overwrite `bytes` of d_buf with a write-only
 * stream, variant 0 = cudaMemsetAsync, 1 = a kernel doing one 16-byte store per thread.
 * bench.py times it next to the table generator as the write-only HBM ceiling. */
int iiv_fill_probe(void* d_buf, size_t bytes, int variant, void* stream);

/* Class constants: MASKED_BITS, MASKED_DOTS, len(BYTE_MASKS), PHASES
 * (screen.py:617-645 HGR, :887-919 DHGR). phases4 receives n_offsets values. */
int iiv_mode_info(int mode, int* masked_bits, int* masked_dots, int* n_offsets,
                  int* phases4);

/* ---- path 1: edit-distance tables (make_data_tables.py) ----------------- */

/* compute_diff_matrix (make_data_tables.py:55-70): 16x16 int()-truncated
 * CIE2000 dE between sRGB triples (colormath 3.0.0 pipeline, FP64, on the
 * device).  h_rgb: 16 x 3 uint8 indexed by nominal colour value
 * (colours.py:27-42); h_lut: 16 x 16 int32.  Synchronises. */
int iiv_lut_cie2000(const uint8_t* h_rgb, int32_t* h_lut);
/* Same, untruncated (diagnostics: distance of each entry to an integer). */
int iiv_lut_cie2000_f64(const uint8_t* h_rgb, double* h_de);

/* Bitmap.to_dots (screen.py:741-789 HGR, :982-990 DHGR) for every masked value:
 * d_dots uint32[n_offsets][2^bits]. */
int iiv_all_dots(int mode, uint32_t* d_dots, void* stream);
/* colours.dots_to_nominal_colour_pixel_values (colours.py:137-148) for every
 * masked value and offset: d_pix uint8[n_offsets][2^bits][masked_dots]. */
int iiv_all_pixel_strings(int mode, uint8_t* d_pix, void* stream);

/* compute_edit_distance (make_data_tables.py:111-174) with edit_distance
 * (:92-108) inlined: fills rows [row_begin,row_end) of the source index i, for
 * every offset, of d_table (the FULL table base pointer).  h_lut: the 16x16
 * substitution costs (compute_substitute_costs, :73-89), each 0..255.
 * Insert/delete are never taken (cost 1e5, :35-36), transposition costs 1. */
int iiv_table_generate(int mode, const int32_t* h_lut, uint16_t* d_table,
                       uint32_t row_begin, uint32_t row_end, int layout,
                       int algo, void* stream);

/* No reference counterpart (host-only, no device needed): the bit windows of the masked
 * value that IIV_ALGO_SPLIT's two halves of edit_distance (make_data_tables.py:92-108)
 * depend on at `offset` -- pixels 0..cut of colours.py:137-148's string are functions of
 * the bits in mask_a only, pixels cut..n-1 of the bits in mask_b only. */
int iiv_table_split_windows(int mode, int offset, int* cut, uint32_t* mask_a,
                            uint32_t* mask_b);

/* edit_distance (make_data_tables.py:92-108) for n_pairs explicit pixel strings
 * of `len` nibble-valued pixels each (d_a, d_b: uint8[n_pairs][len]); d_out
 * int32[n_pairs]. */
int iiv_string_distance(const int32_t* h_lut, const uint8_t* d_a,
                        const uint8_t* d_b, int n_pairs, int len,
                        int32_t* d_out, void* stream);

/* Fused generate + all-gather over NVLink peer memory: rank `rank` of `n_ranks`
 * computes its row block and stores it into every rank's table
 * (h_peer_tables[r] = device pointer of rank r's table, peer-mapped).  If
 * d_multicast_table is non-NULL the block is stored once through the NVSwitch
 * multicast mapping instead.  Callers barrier afterwards. */
int iiv_table_generate_scatter(int mode, const int32_t* h_lut,
                               uint16_t* const* h_peer_tables, int n_ranks,
                               int rank, uint16_t* d_multicast_table,
                               uint32_t row_begin, uint32_t row_end, int layout,
                               void* stream);

/* The device-to-host leg of compute_edit_distance (make_data_tables.py:111-174 returns a host
 * array): rows [row_begin, row_end) of every offset of d_table go to the same place in
 * h_table (both are FULL-table base pointers).  IIV_LAYOUT_TRIANGULAR moves only the columns
 * that can be nonzero (j < i), in `bands` row bands of one 2-D copy each, and does not touch
 * the rest of h_table: the caller passes a buffer whose other bytes are already zero.
 * Asynchronous on `stream` when h_table is page-locked. */
int iiv_table_download(int mode, const uint16_t* d_table, uint16_t* h_table,
                       uint32_t row_begin, uint32_t row_end, int layout, int bands,
                       void* stream);

/* ---- next row N1: the compressed .npz of a table, deflated on the device ---------------
 * make_edit_distance writes np.savez_compressed(..., edit_distance=dist)
 * (make_data_tables.py:186-188).  The table (reference layout, resident in HBM) is cut into
 * blocks of iiv_deflate_block_bytes() bytes; each becomes one dynamic-Huffman deflate block
 * (RFC 1951) plus an empty stored block that byte-aligns it, so the blocks are independent
 * and their concatenation (+ a final empty block) is one raw-deflate stream.  Matches are
 * byte runs that repeat the bytes a whole number of entries back, at a few fixed column
 * distances per mode.  Three stream-ordered steps, the host builds the Huffman codes in
 * between (iivision_b200/deflate.py):
 *   survey  d_hist uint32[n_offsets][320] (286 literal/length + 30 distance counters, from
 *           every sample_every-th block; zeroed here) and d_block_crc uint32[n_blocks], the
 *           CRC-32 of every block; d_crc_ops uint32[7][32] = GF(2) operators appending
 *           256 << k zero bytes to a CRC (zlib's crc32_combine)
 *   encode  d_codes uint32[n_offsets][413]: reversed code | length << 16 for the 286 + 30
 *           symbols, header bit count, header bits (deflate.CodeTable.words());  d_scratch
 *           n_blocks x iiv_deflate_block_stride() bytes, 16-byte aligned; d_sizes
 *           uint32[n_blocks]: compressed bytes of each block (a block that would not shrink
 *           is stored)
 *   gather  packs the blocks back to back: block b to d_out + d_offsets[b] */
size_t iiv_deflate_block_bytes(void);
size_t iiv_deflate_block_stride(void);
int iiv_deflate_survey(int mode, const uint16_t* d_table, uint32_t* d_hist,
                       uint32_t* d_block_crc, const uint32_t* d_crc_ops, int sample_every,
                       void* stream);
int iiv_deflate_encode(int mode, const uint16_t* d_table, const uint32_t* d_codes,
                       uint8_t* d_scratch, uint32_t* d_sizes, void* stream);
int iiv_deflate_gather(const uint8_t* d_scratch, const uint32_t* d_sizes,
                       const int64_t* d_offsets, int n_blocks, uint8_t* d_out, void* stream);

/* Bitmap.edit_distances' in-memory transform (screen.py:358-365):
 * new[y] = old[y] + old[transpose(y)], in place, all offsets. */
int iiv_table_symmetrise(int mode, uint16_t* d_table, void* stream);

/* ---- path 2: per-frame scorer (screen.py) ------------------------------- */

/* Bitmap._pack (screen.py:207-226; _body/_make_header/_make_footer :650-690 HGR,
 * :921-952 DHGR).  d_main/d_aux: uint8[batch][32][256] (d_aux NULL for HGR);
 * d_packed: uint64[batch][32][128].  Strides in bytes between batch items. */
int iiv_pack(int mode, const uint8_t* d_main, const uint8_t* d_aux,
             size_t mem_stride, uint64_t* d_packed, int batch, void* stream);

/* Bitmap.mask_and_shift_data (screen.py:369-378), elementwise over n words. */
int iiv_mask_and_shift(int mode, int byte_offset, const uint64_t* d_in,
                       uint64_t* d_out, size_t n, void* stream);

/* HGRBitmap/DHGRBitmap.masked_update (screen.py:791-816, :992-1007),
 * elementwise over n words (no neighbour fix-up). */
int iiv_masked_update(int mode, int byte_offset, const uint64_t* d_old,
                      uint8_t value, uint64_t* d_new, size_t n, void* stream);

/* Static helpers of the bitmap classes, elementwise over n words:
 *   IIV_PART_HEADER  Bitmap._make_header (screen.py:650-661 HGR, :921-924 DHGR)
 *   IIV_PART_FOOTER  Bitmap._make_footer (screen.py:679-690 HGR, :949-952 DHGR)
 *   IIV_PART_BODY    a packed word with header and footer bits cleared (what
 *                    Bitmap._body returns, screen.py:663-677, :926-947)
 *   IIV_PART_DOUBLE  HGRBitmap._double_pixels (screen.py:710-739), HGR only */
#define IIV_PART_HEADER 0
#define IIV_PART_FOOTER 1
#define IIV_PART_BODY 2
#define IIV_PART_DOUBLE 3
int iiv_column_part(int mode, int part, const uint64_t* d_in, uint64_t* d_out,
                    size_t n, void* stream);

/* Bitmap._fix_column_left (side 0, screen.py:295-306): footer of d_neighbour :=
 * first bits of d_column; _fix_column_right (side 1, :308-320): header of
 * d_neighbour := last bits of d_column.  Elementwise over n words. */
int iiv_fix_column(int mode, int side, const uint64_t* d_neighbour,
                   const uint64_t* d_column, uint64_t* d_out, size_t n, void* stream);

/* Bitmap._fix_array_neighbours (screen.py:322-341) on rows of 128 words. */
int iiv_fix_array_neighbours(int mode, int byte_offset, uint64_t* d_rows,
                             int n_rows, void* stream);

/* Bitmap.diff_weights / _diff_weights (screen.py:400-449): d_out
 * int32[batch][32][256]; source/target uint64[batch][32][128]; d_table is the
 * SYMMETRIC table.  content < 0 means "no content" (the diff_weights call);
 * content >= 0 evaluates every cell as if `content` were stored there. */
int iiv_diff_weights(int mode, int is_aux, const uint64_t* d_source_packed,
                     const uint64_t* d_target_packed, int content,
                     const uint16_t* d_table, int32_t* d_out, int batch,
                     void* stream);

/* The scoring prologue of Video._index_changes (video.py:109-116) for a batch of frames in
 * one launch, fused with Bitmap._pack of each target (screen.py:207-226) and covering both
 * banks of a DHGR frame:
 *   target_packed = _pack(target memory)
 *   diff[bank]    = target.diff_weights(source, bank)          (screen.py:400-449)
 *   diff[bank][SCREEN_HOLES] = 0            if zero_holes       (video.py:111)
 *   priority[bank][diff == 0] = 0; priority[bank] += diff       (video.py:115-116)
 * d_source_packed uint64[batch][32][128] with source_stride words between frames (0 = one
 * source bitmap for every frame); d_target_main / d_target_aux uint8[32][256] per frame,
 * mem_stride bytes apart (aux NULL for HGR); d_target_packed uint64[batch][32][128] (may be
 * NULL); d_diff and d_priority int32[batch][banks][32][256], banks = 1 (HGR) / 2 (DHGR:
 * main, aux); either may be NULL.  d_priority is read and written in place. */
int iiv_score_frames(int mode, const uint64_t* d_source_packed, size_t source_stride,
                     const uint8_t* d_target_main, const uint8_t* d_target_aux,
                     size_t mem_stride, const uint16_t* d_table,
                     uint64_t* d_target_packed, int32_t* d_diff, int32_t* d_priority,
                     int zero_holes, int batch, void* stream);

/* iiv_score_frames without the table.  An edit-distance entry (make_data_tables.py:92-108)
 * is a product of per-pixel 2x2 (min,+) matrices; the product over a segment of pixels is a
 * function of a 4-6 bit window of each masked value, so the entry can be evaluated from small
 * per-segment FACTOR tables (104 KiB per byte offset: a bank's two offsets live in one
 * SM's shared memory) instead of gathered from the 512 MiB / 1 GiB table in HBM.
 *   iiv_score_factors_bytes   size of the factor tables of a mode (all byte offsets, plus a
 *                             16-byte trailer: is the LUT's diagonal zero)
 *   iiv_score_factors         fills d_factors (16-byte aligned) from the 16x16 substitution
 *                             costs compute_substitute_costs gives (make_data_tables.py:73-89);
 *                             the same h_lut as iiv_table_generate
 *   iiv_score_frames_factored iiv_score_frames with d_factors in the place of d_table; every
 *                             output is bit-identical
 *   iiv_score_factor_segments host-only: the segments [p[k], q[k]) of pixels and the bit
 *                             window masks[k] feeding pixels p[k]..min(q[k], n-1) at `offset`
 *                             (arrays of 16), for the CPU tests */
size_t iiv_score_factors_bytes(int mode);
int iiv_score_factors(int mode, const int32_t* h_lut, uint8_t* d_factors, void* stream);
int iiv_score_frames_factored(int mode, const uint64_t* d_source_packed, size_t source_stride,
                              const uint8_t* d_target_main, const uint8_t* d_target_aux,
                              size_t mem_stride, const uint8_t* d_factors,
                              uint64_t* d_target_packed, int32_t* d_diff, int32_t* d_priority,
                              int zero_holes, int batch, void* stream);
int iiv_score_factor_segments(int mode, int offset, int* n_segments, int* p, int* q,
                              uint32_t* masks);

/* Bitmap._diff_weights_page (screen.py:453-494) on n_rows rows of 128 words:
 * d_out int32[n_rows][256]. */
int iiv_diff_weights_page(int mode, int is_aux, const uint64_t* d_source_rows,
                          const uint64_t* d_target_rows, int content,
                          const uint16_t* d_table, int32_t* d_out, int n_rows,
                          void* stream);

/* Bitmap.compute_delta_page (screen.py:525-547): d_out[256] =
 * _diff_weights_page(row, row, is_aux, content) - d_diff_row[256]. */
int iiv_compute_delta_page(int mode, int is_aux,
                           const uint64_t* d_target_packed, int page,
                           int content, const int32_t* d_diff_row,
                           const uint16_t* d_table, int32_t* d_out,
                           void* stream);

/* Every compute_delta_page "new_diff" row of one target frame-bank at once:
 * d_out uint16[batch][32][n_content][256], n_content = 256 (HGR) / 128 (DHGR). */
int iiv_delta_rows(int mode, int is_aux, const uint64_t* d_target_packed,
                   const uint16_t* d_table, uint16_t* d_out, int batch,
                   void* stream);

/* Bitmap.byte_pair_difference (screen.py:383-398) for n (packed word, content)
 * pairs: d_out uint16[n]. */
int iiv_byte_pair_difference(int mode, int byte_offset,
                             const uint64_t* d_old_packed,
                             const uint8_t* d_content, const uint16_t* d_table,
                             uint16_t* d_out, size_t n, void* stream);

/* Bitmap.apply + _fix_scalar_neighbours + MemoryMap.write (screen.py:256-293,
 * :122-125) for a sequence of n stores applied in order to ONE bitmap:
 * h_stores = n x (page, offset, is_aux, value) int32 quadruples. */
int iiv_apply(int mode, uint64_t* d_packed, uint8_t* d_main, uint8_t* d_aux,
              const int32_t* h_stores, int n, void* stream);

/* ---- path 2: encoder (video.py) ------------------------------------------ */

/* Per-clip encoder state blob (Video.__init__, video.py:21-62), one per clip,
 * `iiv_clip_state_bytes()` bytes each, layout given by iiv_clip_state_layout:
 *   [0] packed        uint64[32][128]  Video.pixelmap.packed
 *   [1] main memory   uint8[32][256]   Video.memory_map.page_offset
 *   [2] aux memory    uint8[32][256]   Video.aux_memory_map.page_offset
 *   [3] priority main int32[32][256]   Video.update_priority
 *   [4] priority aux  int32[32][256]   Video.aux_update_priority
 *   [5] mt_numpy      uint32[625]      np.random global MT19937 key + pos
 *   [6] mt_python     uint32[625]      random module MT19937 state + index
 *   [7] flags         int32[8]         [0]=out_of_work main, [1]=out_of_work aux
 */
#define IIV_CLIP_STATE_FIELDS 8
size_t iiv_clip_state_bytes(void);
int iiv_clip_state_layout(size_t* offsets8);

/* Video.encode_frame / _index_changes / _heapify_priorities / _compute_error
 * (video.py:72-301) for n_clips independent clips, one thread block per clip,
 * running the same schedule of n_segments (frame, is_aux, budget) segments
 * (one segment = one encode_frame generator pulled `budget` times).
 *   d_state          n_clips state blobs, state_stride bytes apart
 *   d_target_mem     uint8[n_clips][n_frames][banks][32][256]
 *   d_target_packed  uint64[n_clips][n_frames][32][128] (iiv_pack of the above)
 *   h_segments       int32[n_segments][3]
 *   d_table          symmetric table of the mode
 *   d_opcodes        uint8[n_clips][sum(budget)][8]: page+32, content, o0..o3,
 *                    flag (1 = real opcode, 0 = out-of-work padding), 0
 *   d_seg_info       int64[n_clips][n_segments][8]: real opcodes emitted,
 *                    sum of update_priority before the segment (video.py:90),
 *                    numpy-stream words drawn, python-stream words drawn, then
 *                    tracing counters in SM cycles: score+heapify, opcode loop,
 *                    loop cycles the deciding warp waited for digested heap entries,
 *                    and for MT19937 blocks / room in the store queue
 */
int iiv_encode_clips(int mode, int n_clips, uint8_t* d_state,
                     size_t state_stride, const uint8_t* d_target_mem,
                     const uint64_t* d_target_packed, int n_frames,
                     const int32_t* h_segments, int n_segments,
                     const uint16_t* d_table, uint8_t* d_opcodes,
                     int64_t* d_seg_info, void* stream);

/* Same, for a schedule that is run more than once: the caller keeps a device copy of the
 * segments (d_segments, same contents as h_segments) so the call is a pure kernel launch
 * -- no allocation, no host-to-device copy on the stream.  h_segments is only validated. */
int iiv_encode_clips_planned(int mode, int n_clips, uint8_t* d_state,
                             size_t state_stride, const uint8_t* d_target_mem,
                             const uint64_t* d_target_packed, int n_frames,
                             const int32_t* h_segments, const int32_t* d_segments,
                             int n_segments, const uint16_t* d_table, uint8_t* d_opcodes,
                             int64_t* d_seg_info, void* stream);

/* One encode_frame generator (video.py:72-93) of ONE clip as a single kernel launch, for
 * callers that pull opcodes on the host (the Python facade): the kernel starts from the
 * state blob d_state_in, leaves the state in d_state_out (they may be the same blob), runs
 * `budget` opcodes of (target, is_aux) and writes the opcode records (budget x 8 bytes), the
 * segment info (8 x int64) and the tail of the state blob from field [5] on (both MT19937
 * states and the flags) directly into the PAGE-LOCKED host buffers h_* through their device
 * mappings -- no copies are enqueued (d_opcodes, budget x 8 bytes of device scratch, is where
 * the opcode loop itself writes); `event` (from iiv_event_create, may be NULL) is recorded
 * behind the kernel.  Nothing synchronises: iiv_event_wait does.
 * d_target_mem / d_target_packed describe one frame. */
int iiv_encode_generator(int mode, const uint8_t* d_state_in, uint8_t* d_state_out,
                         const uint8_t* d_target_mem, const uint64_t* d_target_packed,
                         int is_aux, int budget, const uint16_t* d_table,
                         uint8_t* d_opcodes, uint8_t* h_opcodes, int64_t* h_seg_info,
                         uint8_t* h_state_tail, void* event, void* stream);
void* iiv_event_create(void);
int iiv_event_wait(void* event);
int iiv_event_destroy(void* event);

/* ---- next row N2: player byte stream (movie.py, opcodes.py) --------------------- */

/* Movie.emit_stream (movie.py:122-161) with Machine.emit (machine.py:11-25) and the
 * opcodes' emit_command / emit_data (opcodes.py:49-52, 79-89, 116-121, 144-146) for
 * a Header followed by n_ticks tick opcodes, the Acks that close every 2 KiB frame
 * (flipping MAIN/AUX in DHGR), Terminate and the zero padding.
 *   d_opcodes   uint8[n_ticks][8] as written by iiv_encode_clips
 *   d_ticks     uint8[n_ticks]    speaker duty-cycle ticks, 4..66 even (movie.py:104-107)
 *   d_tick_addr uint16[32][32]    op_tick_<4+2i>_page_<32+j> start addresses
 *                                 (opcodes.py:190-217, from player/iivision.dbg)
 *   d_bad       int, set to 1 if a tick or page has no opcode
 * iiv_stream_length gives the byte count (multiple of 2048); iiv_stream_ticks_within
 * the number of ticks Movie.emit_stream emits before max_bytes_out stops it (:133). */
size_t iiv_stream_length(size_t n_ticks, int with_header);
size_t iiv_stream_ticks_within(size_t n_ticks, size_t max_bytes_out);
int iiv_emit_stream(int mode, const uint8_t* d_opcodes, const uint8_t* d_ticks,
                    size_t n_ticks, const uint16_t* d_tick_addr, uint32_t ack_addr,
                    uint32_t terminate_addr, uint8_t* d_out, size_t out_capacity,
                    int* d_bad, void* stream);

/* MT19937 helpers used by the Python facade to keep the process-global
 * generators in step (device-side draws, host-visible state). */
int iiv_mt_draw(uint32_t* d_mt625, uint32_t* d_words, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IIVISION_B200_H_ */

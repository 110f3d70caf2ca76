"""Thin typed wrappers: torch CUDA tensors in, C ABI calls out.

torch is plumbing only (device allocation, streams, pinned host buffers); every
computation happens inside libiivision_b200.so.  Functions here are what the
reference-named facades (make_data_tables.py, screen.py, video.py in this
package) and bench.py call.
"""

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import (ALGO_AUTO, LAYOUT_SYMMETRIC, LAYOUT_TRIANGULAR, MODE_DHGR,
                   MODE_HGR, check, lib)

MODES = {"HGR": MODE_HGR, "DHGR": MODE_DHGR}
MASKED_BITS = {MODE_HGR: 14, MODE_DHGR: 13}
MASKED_DOTS = {MODE_HGR: 18, MODE_DHGR: 10}
NUM_OFFSETS = {MODE_HGR: 2, MODE_DHGR: 4}
NUM_CONTENTS = {MODE_HGR: 256, MODE_DHGR: 128}

STATE_BYTES, STATE_OFFSETS = _lib.clip_state_layout()


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "iivision_b200 needs a CUDA device (B200, sm_100a); there is no "
            "CPU fallback")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor) -> int:
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return t.data_ptr()


def mode_id(mode) -> int:
    if isinstance(mode, str):
        return MODES[mode]
    return int(mode)


def mode_phases(mode):
    """PHASES of the bitmap class (screen.py:645, :919)."""
    return _lib.mode_info(mode_id(mode))[3]


def table_shape(mode):
    m = mode_id(mode)
    return (NUM_OFFSETS[m], 1 << (2 * MASKED_BITS[m]))


# ---- path 1 ------------------------------------------------------------------

def lut_cie2000(rgb16) -> np.ndarray:
    """compute_diff_matrix: uint8[16][3] sRGB -> int32[16][16]."""
    _require_cuda()
    rgb = np.ascontiguousarray(rgb16, dtype=np.uint8).reshape(16, 3)
    out = np.zeros((16, 16), dtype=np.int32)
    check(lib.iiv_lut_cie2000(rgb.ctypes.data, out.ctypes.data))
    return out


def lut_cie2000_float(rgb16) -> np.ndarray:
    _require_cuda()
    rgb = np.ascontiguousarray(rgb16, dtype=np.uint8).reshape(16, 3)
    out = np.zeros((16, 16), dtype=np.float64)
    check(lib.iiv_lut_cie2000_f64(rgb.ctypes.data, out.ctypes.data))
    return out


def all_dots(mode) -> torch.Tensor:
    _require_cuda()
    m = mode_id(mode)
    out = torch.empty((NUM_OFFSETS[m], 1 << MASKED_BITS[m]), dtype=torch.int32,
                      device="cuda")
    check(lib.iiv_all_dots(m, _ptr(out), _stream()))
    return out


def all_pixel_strings(mode) -> torch.Tensor:
    _require_cuda()
    m = mode_id(mode)
    out = torch.empty((NUM_OFFSETS[m], 1 << MASKED_BITS[m], MASKED_DOTS[m]),
                      dtype=torch.uint8, device="cuda")
    check(lib.iiv_all_pixel_strings(m, _ptr(out), _stream()))
    return out


def _lut_arg(lut) -> np.ndarray:
    lut = np.ascontiguousarray(lut, dtype=np.int32)
    if lut.shape != (16, 16):
        raise ValueError("substitution LUT must be 16x16, got %r" % (lut.shape,))
    return lut


def table_generate(mode, lut, layout=LAYOUT_SYMMETRIC, row_begin=0, row_end=None,
                   out: torch.Tensor = None, algo=ALGO_AUTO) -> torch.Tensor:
    """compute_edit_distance on the device.  Returns/fills uint16 (as torch
    int16 bit patterns are avoided: dtype torch.uint16) [n_off, 4**bits]."""
    _require_cuda()
    m = mode_id(mode)
    n_rows = 1 << MASKED_BITS[m]
    if row_end is None:
        row_end = n_rows
    if out is None:
        out = torch.empty(table_shape(m), dtype=torch.uint16, device="cuda")
        if (row_begin, row_end) != (0, n_rows):
            out.zero_()
    if tuple(out.shape) != table_shape(m) or out.dtype != torch.uint16:
        raise ValueError("table must be uint16 %r" % (table_shape(m),))
    lut = _lut_arg(lut)
    check(lib.iiv_table_generate(m, lut.ctypes.data, _ptr(out), row_begin,
                                 row_end, layout, algo, _stream()))
    return out


def fill_probe(buf: torch.Tensor, variant: int = 0) -> None:
    """Write-only stream over ``buf`` (0 = cudaMemsetAsync, 1 = one 16-byte store per
    thread): the ceiling bench.py quotes next to the generator."""
    check(lib.iiv_fill_probe(_ptr(buf), buf.numel() * buf.element_size(), variant, _stream()))


def launches_per_table_generate() -> int:
    """Kernels one iiv_table_generate call launches (the split generator's A / B tables, the
    generator)."""
    return 2


def generator_kernel_name() -> str:
    return "split_kernel"


def table_generate_into(mode, lut, layout, row_begin, row_end,
                        out: torch.Tensor) -> torch.Tensor:
    """Positional form used by parallel.generate_sharded."""
    return table_generate(mode, lut, layout=layout, row_begin=row_begin,
                          row_end=row_end, out=out)


def table_generate_scatter(mode, lut, peer_ptrs, rank, row_begin, row_end,
                           layout=LAYOUT_SYMMETRIC, multicast_ptr=0) -> None:
    _require_cuda()
    m = mode_id(mode)
    lut = _lut_arg(lut)
    n = len(peer_ptrs)
    arr = (ctypes.c_void_p * n)(*[int(p) for p in peer_ptrs])
    check(lib.iiv_table_generate_scatter(
        m, lut.ctypes.data, ctypes.cast(arr, ctypes.c_void_p), n, rank,
        int(multicast_ptr) or None, row_begin, row_end, layout, _stream()))


DOWNLOAD_BANDS = 32


def table_download(mode, table: torch.Tensor, host_ptr: int, row_begin=0, row_end=None,
                   layout=LAYOUT_TRIANGULAR, bands: int = DOWNLOAD_BANDS) -> None:
    """The device-to-host leg of compute_edit_distance: rows [row_begin, row_end) of every
    offset into the host table at ``host_ptr`` (full-table base address, page-locked).  In
    the triangular layout only j < i moves; the rest of the host buffer must already be
    zero.  Stream-ordered; the caller synchronises."""
    m = mode_id(mode)
    if row_end is None:
        row_end = 1 << MASKED_BITS[m]
    if tuple(table.shape) != table_shape(m) or table.dtype != torch.uint16:
        raise ValueError("table must be uint16 %r" % (table_shape(m),))
    check(lib.iiv_table_download(m, _ptr(table), int(host_ptr), row_begin, row_end, layout,
                                 bands, _stream()))


# ---- device-side deflate of a table (next row N1) -----------------------------------------

_deflate_cache = {}


def deflate_table(mode, table: torch.Tensor, sample_every: int = 8, timings: dict = None):
    """Raw-deflate stream of a table resident in HBM (reference layout).  Returns
    (stream uint8 device tensor, block sizes uint32[n_blocks] host, block CRC-32s
    uint32[n_blocks] host, raw bytes per block).  The stream is a sequence of byte-aligned,
    self-contained blocks WITHOUT a final block: append b"\x01\x00\x00\xff\xff".
    ``timings``, when given, receives the device time of each stage in ms (CUDA events)."""
    from . import deflate
    marks = []

    def mark(name):
        if timings is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    m = mode_id(mode)
    if tuple(table.shape) != table_shape(m) or table.dtype != torch.uint16:
        raise ValueError("table must be uint16 %r" % (table_shape(m),))
    block_bytes, stride = int(lib.iiv_deflate_block_bytes()), int(lib.iiv_deflate_block_stride())
    n_off = NUM_OFFSETS[m]
    n_blocks = table.numel() * 2 // block_bytes
    key = ("crc_ops", torch.cuda.current_device())
    if key not in _deflate_cache:
        _deflate_cache[key] = torch.from_numpy(
            deflate.level_operators(256, 7).view(np.int32)).cuda()
    crc_ops = _deflate_cache[key]
    hist = torch.empty((n_off, deflate.HIST_STRIDE), dtype=torch.int32, device="cuda")
    block_crc = torch.empty((n_blocks,), dtype=torch.int32, device="cuda")
    mark("start")
    check(lib.iiv_deflate_survey(m, _ptr(table), _ptr(hist), _ptr(block_crc),
                                 _ptr(crc_ops), int(sample_every), _stream()))
    mark("survey")
    h = hist.cpu().numpy().view(np.uint32)              # synchronises: the codes need it
    codes = np.stack([deflate.CodeTable(h[o]).words() for o in range(n_off)])
    d_codes = torch.from_numpy(codes.view(np.int32)).cuda()
    scratch = torch.empty((n_blocks * stride,), dtype=torch.uint8, device="cuda")
    sizes = torch.empty((n_blocks,), dtype=torch.int32, device="cuda")
    mark("host: huffman codes")
    check(lib.iiv_deflate_encode(m, _ptr(table), _ptr(d_codes), _ptr(scratch),
                                 _ptr(sizes), _stream()))
    mark("encode")
    ends = torch.cumsum(sizes.to(torch.int64), 0)       # bookkeeping of 2^15 sizes
    offsets = (ends - sizes).contiguous()
    total = int(ends[-1].item())
    out = torch.empty((total,), dtype=torch.uint8, device="cuda")
    check(lib.iiv_deflate_gather(_ptr(scratch), _ptr(sizes), _ptr(offsets), n_blocks,
                                 _ptr(out), _stream()))
    mark("scan + gather")
    if timings is not None:
        torch.cuda.synchronize()
        for (_, a), (name, b) in zip(marks, marks[1:]):
            timings[name] = a.elapsed_time(b)
    return (out, sizes.cpu().numpy().view(np.uint32), block_crc.cpu().numpy().view(np.uint32),
            block_bytes)


def table_symmetrise(mode, table: torch.Tensor) -> torch.Tensor:
    m = mode_id(mode)
    if tuple(table.shape) != table_shape(m) or table.dtype != torch.uint16:
        raise ValueError("table must be uint16 %r" % (table_shape(m),))
    check(lib.iiv_table_symmetrise(m, _ptr(table), _stream()))
    return table


def string_distance(lut, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """edit_distance for explicit nibble strings: a, b uint8[n_pairs][len]."""
    _require_cuda()
    lut = _lut_arg(lut)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    if a.shape != b.shape or a.ndim != 2:
        raise ValueError("a and b must be uint8[n_pairs][len] of equal shape")
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty((a.shape[0],), dtype=torch.int32, device="cuda")
    check(lib.iiv_string_distance(lut.ctypes.data, _ptr(da), _ptr(db),
                                  a.shape[0], a.shape[1], _ptr(out), _stream()))
    return out.cpu().numpy()


# ---- path 2: scorer ------------------------------------------------------------

def pack(mode, main: torch.Tensor, aux: torch.Tensor = None) -> torch.Tensor:
    """Bitmap._pack for a batch: uint8[..., 32, 256] -> int64[..., 32, 128]
    (bit pattern of the reference's uint64)."""
    m = mode_id(mode)
    if main.dtype != torch.uint8 or main.shape[-2:] != (32, 256):
        raise ValueError("main memory must be uint8[..., 32, 256]")
    batch = main.numel() // 8192
    out = torch.empty(main.shape[:-2] + (32, 128), dtype=torch.int64,
                      device=main.device)
    if m == MODE_DHGR:
        if aux is None or aux.shape != main.shape or aux.dtype != torch.uint8:
            raise ValueError("DHGR needs aux memory shaped like main")
        aux_ptr = _ptr(aux)
    else:
        aux_ptr = None
    check(lib.iiv_pack(m, _ptr(main), aux_ptr, 8192, _ptr(out), batch, _stream()))
    return out


def pack_strided(mode, base: torch.Tensor, main_off: int, aux_off: int,
                 stride: int, batch: int, out: torch.Tensor) -> torch.Tensor:
    """iiv_pack over interleaved frames: bank b of frame k at
    base + b_off + k * stride (bytes)."""
    m = mode_id(mode)
    p = _ptr(base)
    check(lib.iiv_pack(m, p + main_off, (p + aux_off) if m == MODE_DHGR else None,
                       stride, _ptr(out), batch, _stream()))
    return out


def mask_and_shift(mode, byte_offset, data: torch.Tensor) -> torch.Tensor:
    m = mode_id(mode)
    out = torch.empty_like(data)
    check(lib.iiv_mask_and_shift(m, byte_offset, _ptr(data), _ptr(out),
                                 data.numel(), _stream()))
    return out


def masked_update(mode, byte_offset, old: torch.Tensor, value: int) -> torch.Tensor:
    m = mode_id(mode)
    out = torch.empty_like(old)
    check(lib.iiv_masked_update(m, byte_offset, _ptr(old), int(value) & 0xFF,
                                _ptr(out), old.numel(), _stream()))
    return out


PART_HEADER, PART_FOOTER, PART_BODY, PART_DOUBLE = 0, 1, 2, 3


def column_part(mode, part: int, words: torch.Tensor) -> torch.Tensor:
    """_make_header / _make_footer / body bits / _double_pixels, elementwise."""
    m = mode_id(mode)
    out = torch.empty_like(words)
    check(lib.iiv_column_part(m, part, _ptr(words), _ptr(out), words.numel(), _stream()))
    return out


def fix_column(mode, side: int, neighbour: torch.Tensor, column: torch.Tensor) -> torch.Tensor:
    """_fix_column_left (side 0) / _fix_column_right (side 1), elementwise."""
    m = mode_id(mode)
    if neighbour.shape != column.shape:
        raise ValueError("neighbour and column must have the same shape")
    out = torch.empty_like(neighbour)
    check(lib.iiv_fix_column(m, side, _ptr(neighbour), _ptr(column), _ptr(out),
                             neighbour.numel(), _stream()))
    return out


def fix_array_neighbours(mode, byte_offset, rows: torch.Tensor) -> torch.Tensor:
    m = mode_id(mode)
    if rows.shape[-1] != 128:
        raise ValueError("rows must be [..., 128]")
    check(lib.iiv_fix_array_neighbours(m, byte_offset, _ptr(rows),
                                       rows.numel() // 128, _stream()))
    return rows


def diff_weights(mode, is_aux, source_packed: torch.Tensor,
                 target_packed: torch.Tensor, table: torch.Tensor,
                 content: int = None) -> torch.Tensor:
    """Bitmap.diff_weights for a batch of screens -> int32[..., 32, 256]."""
    m = mode_id(mode)
    if source_packed.shape != target_packed.shape or \
            source_packed.shape[-2:] != (32, 128):
        raise ValueError("packed screens must be [..., 32, 128] and match")
    batch = source_packed.numel() // 4096
    out = torch.empty(source_packed.shape[:-2] + (32, 256), dtype=torch.int32,
                      device=source_packed.device)
    check(lib.iiv_diff_weights(
        m, int(bool(is_aux)), _ptr(source_packed), _ptr(target_packed),
        -1 if content is None else int(content), _ptr(table), _ptr(out), batch,
        _stream()))
    return out


def score_factors(mode, lut) -> torch.Tensor:
    """The factor tables iiv_score_frames_factored evaluates edit-distance entries from
    (uint8 blob, 104 / 106 KiB per byte offset), built from the same 16x16 substitution
    costs as the table."""
    _require_cuda()
    m = mode_id(mode)
    out = torch.empty((lib.iiv_score_factors_bytes(m),), dtype=torch.uint8, device="cuda")
    lut = _lut_arg(lut)
    check(lib.iiv_score_factors(m, lut.ctypes.data, _ptr(out), _stream()))
    return out


def score_factor_segments(mode, offset: int):
    """[(p, q, mask)]: the pixel segments of the factored chain and their bit windows."""
    n = ctypes.c_int()
    p, q, masks = (ctypes.c_int * 16)(), (ctypes.c_int * 16)(), (ctypes.c_uint32 * 16)()
    check(lib.iiv_score_factor_segments(mode_id(mode), int(offset), ctypes.byref(n), p, q, masks))
    return [(p[k], q[k], masks[k]) for k in range(n.value)]


def score_frames(mode, source_packed: torch.Tensor, target_mem: torch.Tensor,
                 table: torch.Tensor = None, priority: torch.Tensor = None,
                 zero_holes: bool = True, want_packed: bool = True, want_diff: bool = True,
                 factors: torch.Tensor = None):
    """The scoring prologue of Video._index_changes (video.py:109-116) for a batch of
    frames in one launch: pack every target, diff_weights of every bank against the
    source bitmap(s), holes zeroed, priorities folded in place.

    source_packed int64[batch, 32, 128] or [32, 128] (one source for all frames);
    target_mem uint8[batch, banks, 32, 256]; priority int32[batch, banks, 32, 256] or
    None.  Returns (target_packed int64[batch, 32, 128] | None, diff int32[batch, banks,
    32, 256] | None).

    The distances come from ``table`` (Bitmap.edit_distances' table, gathered from HBM) or
    from ``factors`` (score_factors(): evaluated from shared memory, no table needed);
    the results are bit-identical."""
    if (table is None) == (factors is None):
        raise ValueError("pass exactly one of table / factors")
    m = mode_id(mode)
    banks = 2 if m == MODE_DHGR else 1
    batch = target_mem.shape[0]
    if target_mem.dtype != torch.uint8 or tuple(target_mem.shape[1:]) != (banks, 32, 256):
        raise ValueError("target_mem must be uint8[batch, %d, 32, 256]" % banks)
    if source_packed.shape == (32, 128):
        stride = 0
    elif tuple(source_packed.shape) == (batch, 32, 128):
        stride = 4096
    else:
        raise ValueError("source_packed must be int64[batch, 32, 128] or [32, 128]")
    if priority is not None and (priority.dtype != torch.int32
                                 or tuple(priority.shape) != (batch, banks, 32, 256)):
        raise ValueError("priority must be int32[batch, %d, 32, 256]" % banks)
    dev = target_mem.device
    tpacked = (torch.empty((batch, 32, 128), dtype=torch.int64, device=dev)
               if want_packed else None)
    diff = (torch.empty((batch, banks, 32, 256), dtype=torch.int32, device=dev)
            if want_diff else None)
    base = _ptr(target_mem)
    fn, dist = ((lib.iiv_score_frames, table) if factors is None
                else (lib.iiv_score_frames_factored, factors))
    check(fn(
        m, _ptr(source_packed), stride, base, (base + 8192) if banks == 2 else None,
        banks * 8192, _ptr(dist), _ptr(tpacked) if want_packed else None,
        _ptr(diff) if want_diff else None, _ptr(priority) if priority is not None else None,
        int(bool(zero_holes)), batch, _stream()))
    return tpacked, diff


def diff_weights_page(mode, is_aux, source_rows: torch.Tensor,
                      target_rows: torch.Tensor, table: torch.Tensor,
                      content: int = None) -> torch.Tensor:
    m = mode_id(mode)
    n_rows = source_rows.numel() // 128
    out = torch.empty(source_rows.shape[:-1] + (256,), dtype=torch.int32,
                      device=source_rows.device)
    check(lib.iiv_diff_weights_page(
        m, int(bool(is_aux)), _ptr(source_rows), _ptr(target_rows),
        -1 if content is None else int(content), _ptr(table), _ptr(out), n_rows,
        _stream()))
    return out


def compute_delta_page(mode, is_aux, target_packed: torch.Tensor, page: int,
                       content: int, diff_row: torch.Tensor,
                       table: torch.Tensor) -> torch.Tensor:
    m = mode_id(mode)
    out = torch.empty((256,), dtype=torch.int32, device=target_packed.device)
    check(lib.iiv_compute_delta_page(
        m, int(bool(is_aux)), _ptr(target_packed), int(page), int(content),
        _ptr(diff_row), _ptr(table), _ptr(out), _stream()))
    return out


def delta_rows(mode, is_aux, target_packed: torch.Tensor,
               table: torch.Tensor) -> torch.Tensor:
    """All compute_delta_page new-diff rows: uint16[..., 32, n_content, 256]."""
    m = mode_id(mode)
    batch = target_packed.numel() // 4096
    out = torch.empty(target_packed.shape[:-2] + (32, NUM_CONTENTS[m], 256),
                      dtype=torch.uint16, device=target_packed.device)
    check(lib.iiv_delta_rows(m, int(bool(is_aux)), _ptr(target_packed),
                             _ptr(table), _ptr(out), batch, _stream()))
    return out


def byte_pair_difference(mode, byte_offset, old_packed: torch.Tensor,
                         content: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    m = mode_id(mode)
    out = torch.empty(old_packed.shape, dtype=torch.uint16,
                      device=old_packed.device)
    check(lib.iiv_byte_pair_difference(
        m, byte_offset, _ptr(old_packed), _ptr(content), _ptr(table), _ptr(out),
        old_packed.numel(), _stream()))
    return out


def apply_stores(mode, packed: torch.Tensor, main: torch.Tensor,
                 aux: torch.Tensor, stores) -> None:
    """Bitmap.apply for a sequence of (page, offset, is_aux, value) stores."""
    m = mode_id(mode)
    st = np.ascontiguousarray(stores, dtype=np.int32).reshape(-1, 4)
    check(lib.iiv_apply(m, _ptr(packed), _ptr(main),
                        _ptr(aux) if aux is not None else None,
                        st.ctypes.data, st.shape[0], _stream()))


# ---- path 2: encoder -----------------------------------------------------------

def new_clip_states(n_clips: int) -> torch.Tensor:
    """Zeroed Video.__init__ state blobs; RNG fields must be set by the caller."""
    _require_cuda()
    return torch.zeros((n_clips, STATE_BYTES), dtype=torch.uint8, device="cuda")


def state_field(states: torch.Tensor, field: int, dtype, shape) -> torch.Tensor:
    """View of one field of every clip state: [n_clips, *shape]."""
    off = STATE_OFFSETS[field]
    nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    return states[:, off:off + nbytes].view(dtype).view((states.shape[0],) + tuple(shape))


F_PACKED, F_MAIN, F_AUX, F_PRIO_MAIN, F_PRIO_AUX, F_MT_NP, F_MT_PY, F_FLAGS = range(8)


def seed_clip_states(states: torch.Tensor, seeds) -> torch.Tensor:
    """``random.seed(s); np.random.seed(s)`` for every clip: both MT19937 fields of clip k
    are set to the states those calls leave behind (init_by_array / init_genrand)."""
    import random
    n = states.shape[0]
    seeds = list(seeds)
    if len(seeds) != n:
        raise ValueError("need one seed per clip")
    mt_np = np.zeros((n, 640), dtype=np.uint32)
    mt_py = np.zeros((n, 640), dtype=np.uint32)
    for k, s in enumerate(seeds):
        mt_py[k, :625] = mt_from_python(random.Random(int(s)).getstate())
        mt_np[k, :625] = mt_from_numpy(np.random.RandomState(int(s)).get_state())
    state_field(states, F_MT_NP, torch.int32, (640,)).copy_(
        torch.from_numpy(mt_np.view(np.int32)).to(states.device))
    state_field(states, F_MT_PY, torch.int32, (640,)).copy_(
        torch.from_numpy(mt_py.view(np.int32)).to(states.device))
    return states


class SegmentPlan:
    """A (frame, is_aux, budget) schedule kept on the device as well, for encode_clips
    calls that repeat it: no allocation or host-to-device copy per call."""

    def __init__(self, segments):
        self.host = np.ascontiguousarray(segments, dtype=np.int32).reshape(-1, 3)
        self.device = torch.from_numpy(self.host).cuda()
        self.total = int(self.host[:, 2].sum())

    def __len__(self):
        return self.host.shape[0]


def encode_clips(mode, states: torch.Tensor, target_mem: torch.Tensor,
                 target_packed: torch.Tensor, segments, table: torch.Tensor,
                 opcodes: torch.Tensor = None, seg_info: torch.Tensor = None):
    """Runs the (frame, is_aux, budget) schedule for every clip.

    target_mem uint8[n_clips, n_frames, banks, 32, 256]; target_packed
    int64[n_clips, n_frames, 32, 128].  Returns (opcodes uint8[n_clips,
    total_budget, 8], seg_info int64[n_clips, n_segments, 8])."""
    m = mode_id(mode)
    plan = segments if isinstance(segments, SegmentPlan) else None
    segs = plan.host if plan else np.ascontiguousarray(segments, dtype=np.int32).reshape(-1, 3)
    n_clips, n_frames = target_mem.shape[0], target_mem.shape[1]
    banks = 2 if m == MODE_DHGR else 1
    if target_mem.shape != (n_clips, n_frames, banks, 32, 256):
        raise ValueError("target_mem must be uint8[n_clips, n_frames, %d, 32, 256]" % banks)
    if target_packed.shape != (n_clips, n_frames, 32, 128):
        raise ValueError("target_packed must be int64[n_clips, n_frames, 32, 128]")
    if states.shape != (n_clips, STATE_BYTES):
        raise ValueError("states must be uint8[n_clips, %d]" % STATE_BYTES)
    total = int(segs[:, 2].sum())
    if opcodes is None:
        opcodes = torch.empty((n_clips, total, 8), dtype=torch.uint8, device="cuda")
    if seg_info is None:
        seg_info = torch.zeros((n_clips, segs.shape[0], 8), dtype=torch.int64,
                               device="cuda")
    if plan is not None:
        check(lib.iiv_encode_clips_planned(
            m, n_clips, _ptr(states), STATE_BYTES, _ptr(target_mem), _ptr(target_packed),
            n_frames, segs.ctypes.data, _ptr(plan.device), segs.shape[0], _ptr(table),
            opcodes.data_ptr() or None, _ptr(seg_info), _stream()))
    else:
        check(lib.iiv_encode_clips(
            m, n_clips, _ptr(states), STATE_BYTES, _ptr(target_mem),
            _ptr(target_packed), n_frames, segs.ctypes.data, segs.shape[0],
            _ptr(table), opcodes.data_ptr() or None, _ptr(seg_info), _stream()))
    return opcodes, seg_info


def mt_draw(mt625: torch.Tensor, n: int) -> torch.Tensor:
    words = torch.empty((max(n, 1),), dtype=torch.int32, device="cuda")
    check(lib.iiv_mt_draw(_ptr(mt625), _ptr(words), n, _stream()))
    return words[:n]


# ---- host <-> device RNG state plumbing ---------------------------------------------

def mt_from_python(state) -> np.ndarray:
    """random.getstate() -> uint32[625] (624 words + index)."""
    version, internal, _ = state
    if version != 3 or len(internal) != 625:
        raise ValueError("unexpected random.getstate() layout")
    return np.array(internal, dtype=np.uint32)


def mt_to_python(words625: np.ndarray):
    return (3, tuple(np.asarray(words625).tolist()), None)


def mt_from_numpy(state) -> np.ndarray:
    """np.random.get_state() (legacy tuple) -> uint32[625]."""
    name, key, pos = state[0], state[1], state[2]
    if name != "MT19937":
        raise ValueError("legacy numpy RNG is not MT19937")
    out = np.empty(625, dtype=np.uint32)
    out[:624] = key
    out[624] = pos
    return out


def mt_to_numpy(words625: np.ndarray):
    return ("MT19937", np.array(words625[:624], dtype=np.uint32),
            int(words625[624]), 0, 0.0)

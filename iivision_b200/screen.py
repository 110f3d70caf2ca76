"""Representations of the Apple II (D)HGR display (reference transcoder/screen.py),
with every array computation done by the CUDA library.

Same names, signatures and error behaviour as the reference module, so that code
written against ``screen`` keeps working (``import iivision_b200.screen as
screen``).  What differs is where the work happens:

* ``Bitmap.packed`` stays a host ``uint64[32][128]`` array (reference code reads
  and writes it directly), but ``_pack``, ``mask_and_shift_data``,
  ``masked_update``, ``_fix_array_neighbours``, ``apply``, ``diff_weights``,
  ``_diff_weights_page``, ``compute_delta_page`` and ``byte_pair_difference`` are
  C-ABI calls (include/iivision_b200.h); host arrays are staged to the device and
  back around each call.  There is no numpy implementation to fall back to.
* ``Bitmap.edit_distances`` returns the symmetric table like the reference
  (screen.py:343-367) but builds it on the GPU: from the reference's ``.npz`` when
  it exists (load + device-side transpose-add), otherwise generated in place
  (make_data_tables.compute_edit_distance on the device, milliseconds).  The
  device copy stays resident for the scoring kernels
  (``Bitmap.edit_distances_device``).

For throughput use ``iivision_b200.video.Video`` (whole encode segments in one
kernel) or ``iivision_b200.ops`` (batched, device-resident); the per-call methods
here exist for drop-in compatibility.
"""

import functools
import os
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from . import npz_io
from . import ops
from . import palette as pal

IntOrArray = Union[np.uint64, np.ndarray]


def y_to_base_addr(y: int, page: int = 0) -> int:
    """Maps y coordinate to base address on given screen page.

    The hi-res interleave: the 8 lines of a character row sit 1 KiB apart, the 8 character
    rows of a third of the screen 128 bytes apart, the three thirds 40 bytes apart (which
    is what leaves 8 unused bytes behind every third row: the screen holes)."""
    line, row, third = y & 7, (y >> 3) & 7, y >> 6
    return ((page + 1) << 13) + (line << 10) + (row << 7) + 40 * third


Y_TO_BASE_ADDR = [[y_to_base_addr(y, p) for y in range(192)] for p in (0, 1)]


def _address_tables():
    """The (y, x) <-> (page, offset) tables of reference screen.py:26-69, built from the
    address of every visible byte of page 1 at once."""
    addr = np.asarray(Y_TO_BASE_ADDR[0])[:, None] + np.arange(40)[None, :]      # [192][40]
    page, offset = (addr >> 8) - 32, addr & 255
    ys, xs = np.broadcast_arrays(np.arange(192)[:, None], np.arange(40)[None, :])
    to_x = np.zeros((32, 256), dtype=np.uint8)
    to_y = np.zeros((32, 256), dtype=np.uint8)
    holes = np.ones((32, 256), dtype=np.bool_)
    to_x[page, offset] = xs
    to_y[page, offset] = ys
    holes[page, offset] = False
    coords = {}
    for p in (0, 1):
        base = np.asarray(Y_TO_BASE_ADDR[p])[:, None] + np.arange(40)[None, :]
        coords.update({int(a): (p, int(y), int(x))
                       for a, y, x in zip(base.ravel(), ys.ravel(), xs.ravel())})
    return to_x, to_y, page.astype(np.uint8), offset.astype(np.uint8), holes, coords


(PAGE_OFFSET_TO_X, PAGE_OFFSET_TO_Y, X_Y_TO_PAGE, X_Y_TO_OFFSET, SCREEN_HOLES,
 ADDR_TO_COORDS) = _address_tables()


def _checked_screen_page(screen_page: int) -> int:
    if screen_page not in (1, 2):
        raise ValueError("Screen page out of bounds: %d" % screen_page)
    return screen_page


def _bytes_or_zeros(data, shape) -> np.ndarray:
    """The caller's array (adopted, not copied, as the reference does) or a blank screen."""
    if data is None:
        return np.zeros(shape, dtype=np.uint8)
    if data.shape != shape:
        raise ValueError("Unexpected shape: %r" % (data.shape,))
    return data


class FlatMemoryMap:
    """Linear 8K representation of HGR screen memory."""

    def __init__(self, screen_page: int, data: np.array = None):
        self.screen_page = _checked_screen_page(screen_page)
        self._addr_start = 8192 * self.screen_page
        self._addr_end = self._addr_start + 8191
        self.data = _bytes_or_zeros(data, (8192,))

    def to_memory_map(self):
        return MemoryMap(self.screen_page, self.data.reshape((32, 256)))

    def write(self, addr: int, val: int) -> None:
        """Updates screen image to set 0xaddr = val (including screen holes)"""
        if not self._addr_start <= addr <= self._addr_end:
            raise ValueError("Address out of range: 0x%04x" % addr)
        self.data[addr - self._addr_start] = val


class MemoryMap:
    """Page/offset-structured representation of HGR screen memory."""

    def __init__(self, screen_page: int, page_offset: np.array = None):
        self.screen_page = _checked_screen_page(screen_page)
        self._page_start = 32 * screen_page
        self.page_offset = _bytes_or_zeros(page_offset, (32, 256))

    def to_flat_memory_map(self) -> FlatMemoryMap:
        return FlatMemoryMap(self.screen_page, self.page_offset.reshape(8192))

    def write(self, page: int, offset: int, val: int) -> None:
        """Updates screen image to set (page, offset)=val (inc. screen holes)"""
        # page may be 0..31 (negative index after the subtraction lands on the
        # same row) or 32..63, exactly like the reference (screen.py:122-125)
        self.page_offset[page - self._page_start][offset] = val


# ---- host <-> device staging ------------------------------------------------------

def _h2d_u64(a) -> torch.Tensor:
    a = np.require(a, dtype=np.uint64, requirements=["C", "W"])   # copies read-only input
    return torch.from_numpy(a.view(np.int64)).cuda()


def _d2h_u64(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint64)


def _h2d_u8(a) -> torch.Tensor:
    return torch.from_numpy(np.require(a, dtype=np.uint8, requirements=["C", "W"])).cuda()


class Bitmap:
    """Packed bitmap representation of (D)HGR screen memory."""

    NAME = None  # type: str
    MODE = None  # type: int   (IIV_MODE_*)

    HEADER_BITS = None  # type: np.uint64
    BODY_BITS = None  # type: np.uint64
    FOOTER_BITS = None  # type: np.uint64
    MASKED_BITS = None  # type: np.uint64
    MASKED_DOTS = None  # type: np.uint64
    BYTE_MASKS = None  # type: List[np.uint64]
    BYTE_SHIFTS = None  # type: List[np.uint64]
    PHASES = None  # type: List[int]

    # Where edit_distances looks for the reference's table files (screen.py:348)
    DATA_DIR = "transcoder/data"

    def __init__(self, palette: pal.Palette, main_memory: MemoryMap,
                 aux_memory: Optional[MemoryMap]):
        self.palette = palette
        self.main_memory = main_memory
        self.aux_memory = aux_memory
        self.PACKED_BITS = self.HEADER_BITS + self.BODY_BITS + self.FOOTER_BITS
        self.SCREEN_BYTES = np.uint64(len(self.BYTE_MASKS))
        self._packed = np.empty(shape=(32, 128), dtype=np.uint64)
        self._pack()

    # -- packing -----------------------------------------------------------------
    @classmethod
    def _part(cls, part: int, col: IntOrArray) -> IntOrArray:
        if cls.MODE is None:
            raise NotImplementedError
        scalar = np.ndim(col) == 0
        arr = np.atleast_1d(np.asarray(col, dtype=np.uint64))
        out = _d2h_u64(ops.column_part(cls.MODE, part, _h2d_u64(arr))).reshape(arr.shape)
        return np.uint64(out[0]) if scalar else out

    @classmethod
    def _make_header(cls, col: IntOrArray) -> IntOrArray:
        """Extract values to use as header of next column."""
        return cls._part(ops.PART_HEADER, col)

    @classmethod
    def _make_footer(cls, col: IntOrArray) -> IntOrArray:
        """Extract values to use as footer of previous column."""
        return cls._part(ops.PART_FOOTER, col)

    def _body(self) -> np.ndarray:
        """Pack related screen bytes into an efficient representation (no header /
        footer bits)."""
        main = _h2d_u8(self.main_memory.page_offset)
        aux = _h2d_u8(self.aux_memory.page_offset) if self.aux_memory is not None \
            and self.MODE == ops.MODE_DHGR else None
        return _d2h_u64(ops.column_part(self.MODE, ops.PART_BODY,
                                        ops.pack(self.MODE, main, aux)))

    def _fix_column_left(self, column_left: IntOrArray, column: IntOrArray) -> IntOrArray:
        """Patch up the footer of the column to the left."""
        return self._fix_column(0, column_left, column)

    def _fix_column_right(self, column_right: IntOrArray, column: IntOrArray) -> IntOrArray:
        """Patch up the header of the column to the right."""
        return self._fix_column(1, column_right, column)

    def _fix_column(self, side, neighbour, column):
        scalar = np.ndim(neighbour) == 0
        a = np.atleast_1d(np.asarray(neighbour, dtype=np.uint64))
        b = np.broadcast_to(np.asarray(column, dtype=np.uint64), a.shape)
        out = _d2h_u64(ops.fix_column(self.MODE, side, _h2d_u64(a), _h2d_u64(b))).reshape(a.shape)
        return np.uint64(out[0]) if scalar else out

    def _pack(self) -> None:
        """Pack MemoryMap into efficient representation for diffing
        (screen.py:207-226 -> iiv_pack).  The packing runs on the device and nothing waits
        for it: ``packed`` is fetched when it is first read.  The device copies of the
        memory maps and of the packed words are kept, with the host values they were made
        from, so that video.Video can score against this bitmap without uploading it."""
        dhgr = self.aux_memory is not None and self.MODE == ops.MODE_DHGR
        banks = [np.require(self.main_memory.page_offset, dtype=np.uint8, requirements=["C"])]
        if dhgr:
            banks.append(np.require(self.aux_memory.page_offset, dtype=np.uint8,
                                    requirements=["C"]))
        host = np.stack(banks)
        tmem = torch.from_numpy(host).cuda()
        dev = ops.pack(self.MODE, tmem[0], tmem[1] if dhgr else None)
        self._packed = None
        self._device_copy = (host, None, tmem, dev)

    @property
    def packed(self) -> np.ndarray:
        """uint64[32][128] (screen.py:186-188), brought from the device on first use."""
        if self._packed is None:
            host, _, tmem, dev = self._device_copy
            self._packed = _d2h_u64(dev)
            self._device_copy = (host, self._packed.copy(), tmem, dev)
        return self._packed

    @packed.setter
    def packed(self, value) -> None:
        self._packed = value

    @classmethod
    def masked_update(cls, byte_offset: int, old_value: IntOrArray,
                      new_value: np.uint8) -> IntOrArray:
        """Update int/array to store new value at byte_offset in every entry.
        Does not patch up headers/footers of neighbouring columns."""
        if cls.MODE is None:
            raise NotImplementedError
        scalar = np.ndim(old_value) == 0
        arr = np.atleast_1d(np.asarray(old_value, dtype=np.uint64))
        out = _d2h_u64(ops.masked_update(cls.MODE, int(byte_offset), _h2d_u64(arr),
                                         int(new_value) & 0xFF)).reshape(arr.shape)
        return np.uint64(out[0]) if scalar else out

    @staticmethod
    def byte_offset(page_offset: int, is_aux: bool) -> int:
        raise NotImplementedError

    @staticmethod
    def _byte_offsets(is_aux: bool) -> Tuple[int, int]:
        raise NotImplementedError

    @classmethod
    def to_dots(cls, masked_val: int, byte_offset: int) -> int:
        """Convert masked representation to bit sequence of display dots."""
        if cls.MODE is None:
            raise NotImplementedError
        return int(cls._all_dots()[byte_offset, masked_val])

    @classmethod
    @functools.lru_cache(None)
    def _all_dots(cls) -> np.ndarray:
        return ops.all_dots(cls.MODE).cpu().numpy().view(np.uint32)

    @classmethod
    def mask_and_shift_data(cls, data: IntOrArray, byte_offset: int) -> IntOrArray:
        """Masks and shifts packed data into the MASKED_BITS range."""
        scalar = np.ndim(data) == 0
        arr = np.atleast_1d(np.asarray(data, dtype=np.uint64))
        out = _d2h_u64(ops.mask_and_shift(cls.MODE, int(byte_offset),
                                          _h2d_u64(arr))).reshape(arr.shape)
        return np.uint64(out[0]) if scalar else out

    # -- single-byte update ------------------------------------------------------------
    def apply(self, page: int, offset: int, is_aux: bool, value: np.uint8) -> None:
        """Update packed representation of changing main/aux memory
        (screen.py:256-293 -> iiv_apply)."""
        self.apply_many([(page, offset, is_aux, value)])

    def apply_many(self, stores) -> None:
        """apply() for a sequence of (page, offset, is_aux, value), in order."""
        if any(bool(s[2]) for s in stores) and self.aux_memory is None:
            raise ValueError("aux store on a bitmap without aux memory")
        packed = _h2d_u64(self.packed)
        main = _h2d_u8(self.main_memory.page_offset)
        aux = _h2d_u8(self.aux_memory.page_offset) if self.aux_memory is not None else None
        ops.apply_stores(self.MODE, packed, main, aux,
                         [(int(p) % 32, int(o), int(bool(a)), int(v)) for p, o, a, v in stores])
        self.packed[...] = _d2h_u64(packed)
        self.main_memory.page_offset[...] = main.cpu().numpy()
        if aux is not None:
            self.aux_memory.page_offset[...] = aux.cpu().numpy()

    def _fix_array_neighbours(self, ary: np.ndarray, byte_offset: int) -> None:
        """Fix up column headers/footers for all array entries, in place."""
        rows = _h2d_u64(ary.reshape(-1, 128))
        ops.fix_array_neighbours(self.MODE, int(byte_offset), rows)
        ary[...] = _d2h_u64(rows).reshape(ary.shape)

    # -- tables ------------------------------------------------------------------------
    @classmethod
    def _table_path(cls, palette_id: pal.Palette) -> str:
        return "%s/%s_palette_%d_edit_distance.npz" % (
            cls.DATA_DIR, cls.NAME, palette_id.value)

    @classmethod
    @functools.lru_cache(None)
    def edit_distances_device(cls, palette_id: pal.Palette) -> torch.Tensor:
        """Symmetric uint16[n_offsets, 4**bits] table resident in HBM."""
        path = cls._table_path(palette_id)
        if os.path.exists(path):
            # inflated on all host cores into page-locked memory when the file carries our
            # writer's piece index, through np.load otherwise (npz_io.py)
            tri = npz_io.load_member(
                path, "edit_distance",
                alloc=lambda n: torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy())
            want = ops.table_shape(cls.MODE)
            if tri.shape != want or tri.dtype != np.uint16:
                raise ValueError("%s: expected uint16 %r" % (path, want))
            table = torch.from_numpy(tri.view(np.int16)).cuda().view(torch.uint16)
            return ops.table_symmetrise(cls.MODE, table)   # screen.py:358-365
        lut = ops.lut_cie2000(pal.PALETTES[palette_id].rgb_by_value())
        return ops.table_generate(cls.MODE, lut, layout=ops.LAYOUT_SYMMETRIC)

    @classmethod
    @functools.lru_cache(None)
    def edit_distances(cls, palette_id: pal.Palette) -> np.ndarray:
        """Load edit distance matrices for masked, shifted byte values."""
        return cls.edit_distances_device(palette_id).cpu().numpy()

    # -- scoring -------------------------------------------------------------------------
    def byte_pair_difference(self, byte_offset: int, old_packed: np.uint64,
                             content: np.uint8) -> np.uint16:
        """Compute effect of storing a new content byte within packed data."""
        out = ops.byte_pair_difference(
            self.MODE, int(byte_offset), _h2d_u64([old_packed]),
            _h2d_u8([int(content) & 0xFF]), self.edit_distances_device(self.palette))
        return np.uint16(out.cpu().numpy().view(np.uint16)[0])

    def diff_weights(self, source: "Bitmap", is_aux: bool) -> np.ndarray:
        """Compute edit distance matrix from source bitmap."""
        return self._diff_weights(source.packed, is_aux)

    def _diff_weights(self, source_packed: np.ndarray, is_aux: bool,
                      content: np.uint8 = None) -> np.ndarray:
        out = ops.diff_weights(
            self.MODE, is_aux, _h2d_u64(source_packed), _h2d_u64(self.packed),
            self.edit_distances_device(self.palette),
            None if content is None else int(content))
        return out.cpu().numpy()

    def _diff_weights_page(self, source_packed: np.ndarray, target_packed: np.ndarray,
                           is_aux: bool, content: np.uint8 = None) -> np.ndarray:
        out = ops.diff_weights_page(
            self.MODE, is_aux, _h2d_u64(np.reshape(source_packed, (128,))),
            _h2d_u64(np.reshape(target_packed, (128,))),
            self.edit_distances_device(self.palette),
            None if content is None else int(content))
        return out.cpu().numpy()

    def compute_delta_page(self, page: int, content: int, diff_weights: np.ndarray,
                           is_aux: bool) -> np.ndarray:
        """Compute which content stores introduce the least additional error:
        new diff of storing content at every offset minus the previous weights."""
        row = torch.from_numpy(np.ascontiguousarray(diff_weights, dtype=np.int32)).cuda()
        out = ops.compute_delta_page(
            self.MODE, is_aux, _h2d_u64(self.packed), int(page), int(content), row,
            self.edit_distances_device(self.palette))
        return out.cpu().numpy()

    def _check_consistency(self):
        """Sanity check that headers and footers are consistent with a repack."""
        main = _h2d_u8(self.main_memory.page_offset)
        aux = _h2d_u8(self.aux_memory.page_offset) if self.aux_memory is not None \
            and self.MODE == ops.MODE_DHGR else None
        assert np.array_equal(_d2h_u64(ops.pack(self.MODE, main, aux)), self.packed)


class HGRBitmap(Bitmap):
    """Packed bitmap representation of HGR screen memory: 22-bit words
    ffFbbbbbbbBAaaaaaaaHhh for each pair of screen bytes (screen.py:550-645)."""

    NAME = 'HGR'
    MODE = ops.MODE_HGR

    HEADER_BITS = np.uint64(3)
    BODY_BITS = np.uint64(16)
    FOOTER_BITS = np.uint64(3)
    MASKED_BITS = np.uint64(14)
    MASKED_DOTS = np.uint64(18)

    BYTE_MASKS = [np.uint64(0x3FFF), np.uint64(0x3FFF << 8)]
    BYTE_SHIFTS = [np.uint64(0), np.uint64(8)]
    PHASES = [1, 3]

    def __init__(self, palette: pal.Palette, main_memory: MemoryMap):
        super(HGRBitmap, self).__init__(palette, main_memory, None)

    @classmethod
    def _double_pixels(cls, int7: int) -> int:
        """Each bit 0..5 becomes two dots, bit 6 three (screen.py:710-739)."""
        return int(cls._part(ops.PART_DOUBLE, np.uint64(int7)))

    @staticmethod
    @functools.lru_cache(None)
    def byte_offset(page_offset: int, is_aux: bool) -> int:
        """Returns 0..1 offset in packed representation for page_offset."""
        assert not is_aux
        return int(page_offset) % 2

    @staticmethod
    @functools.lru_cache(None)
    def _byte_offsets(is_aux: bool) -> Tuple[int, int]:
        assert not is_aux
        return 0, 1


class DHGRBitmap(Bitmap):
    """Packed bitmap representation of DHGR screen memory: 34-bit words, 3-bit
    header, four 7-bit bytes (aux/main interleaved), 3-bit footer
    (screen.py:819-919)."""

    NAME = 'DHGR'
    MODE = ops.MODE_DHGR

    HEADER_BITS = np.uint64(3)
    BODY_BITS = np.uint64(28)
    FOOTER_BITS = np.uint64(3)
    MASKED_BITS = np.uint64(13)
    MASKED_DOTS = np.uint64(10)

    BYTE_MASKS = [np.uint64(0x1FFF << (7 * k)) for k in range(4)]
    BYTE_SHIFTS = [np.uint64(7 * k) for k in range(4)]
    PHASES = [1, 0, 3, 2]

    @staticmethod
    @functools.lru_cache(None)
    def byte_offset(page_offset: int, is_aux: bool) -> int:
        """Returns 0..3 packed byte offset for a given page_offset and is_aux"""
        odd = int(page_offset) % 2 == 1
        if is_aux:
            return 2 if odd else 0
        return 3 if odd else 1

    @staticmethod
    @functools.lru_cache(None)
    def _byte_offsets(is_aux: bool) -> Tuple[int, int]:
        return (0, 2) if is_aux else (1, 3)

"""CPU restatement (numpy + heapq) of the reference's per-frame scorer/encoder.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by iivision_b200.

PARITY STATUS: pinned.  tests/test_oracle_vs_reference.py runs this file against
the UNMODIFIED reference modules (oracle/ref_harness.py) on random screens and
whole encode runs, and tests/golden/ holds outputs generated from the reference
itself (oracle/make_golden.py) including the reference's own unit-test literals
(screen_test.py, video_test.py:28-30, 48-51, 63-66).

It keeps the reference's cost structure on purpose (whole-array numpy calls per
opcode, heapq, Python-level RNG draws) so that timing it is a fair stand-in for
the reference's single-threaded numpy path where the reference cannot travel.

Reference lines restated:
  layout / hole map          screen.py:16-69
  pack                       screen.py:207-226, 650-690 (HGR), 921-952 (DHGR)
  mask_shift                 screen.py:369-378
  masked_update              screen.py:791-816 (HGR), 992-1007 (DHGR)
  byte_offset(s)             screen.py:692-708 (HGR), 954-980 (DHGR)
  apply / neighbour fix-ups  screen.py:256-341
  diff_weights(_page)        screen.py:400-494
  compute_delta_page         screen.py:525-547
  byte_pair_difference       screen.py:383-398
  Encoder                    video.py:21-62, 72-301
"""

import heapq
import random as _pyrandom

import numpy as np

U64 = np.uint64


def _hole_mask() -> np.ndarray:
    """screen.py:16-69: (page, offset) cells not backing any of the 192x40
    screen bytes (offsets 120..127 and 248..255 of every page)."""
    holes = np.ones((32, 256), dtype=bool)
    for y in range(192):
        a, d = divmod(y, 64)
        b, c = divmod(d, 8)
        base = 8192 + 1024 * c + 128 * b + 40 * a
        page, off = divmod(base, 256)
        holes[page - 32, off:off + 40] = False
    return holes


SCREEN_HOLES = _hole_mask()


def xy_to_page_offset(x_byte: int, y: int):
    a, d = divmod(y, 64)
    b, c = divmod(d, 8)
    addr = 8192 + 1024 * c + 128 * b + 40 * a + x_byte
    return (addr >> 8) - 32, addr & 0xFF


class ModeSpec:
    """Per-mode constants and bit plumbing (HGRBitmap / DHGRBitmap class
    attributes and static methods)."""

    def __init__(self, name):
        self.name = name
        if name == "HGR":
            self.header_bits, self.body_bits, self.footer_bits = 3, 16, 3
            self.masked_bits = 14
            self.masks = [U64(0x3FFF), U64(0x3FFF << 8)]
            self.shifts = [U64(0), U64(8)]
        elif name == "DHGR":
            self.header_bits, self.body_bits, self.footer_bits = 3, 28, 3
            self.masked_bits = 13
            self.masks = [U64(0x1FFF << (7 * k)) for k in range(4)]
            self.shifts = [U64(7 * k) for k in range(4)]
        else:
            raise ValueError(name)
        self.n_offsets = len(self.masks)
        self.keep_low = U64((1 << (self.header_bits + self.body_bits)) - 1)
        self.keep_high = U64(
            ((1 << (self.body_bits + self.footer_bits)) - 1)
            << self.header_bits)

    # -- packing -----------------------------------------------------------
    def body(self, main, aux):
        if self.name == "HGR":
            even = main[:, 0::2].astype(U64)
            odd = main[:, 1::2].astype(U64)
            return ((even << U64(3)) + ((odd & U64(0x7F)) << U64(12))
                    + ((odd & U64(0x80)) << U64(4)))
        a = (aux & 0x7F).astype(U64)
        m = (main & 0x7F).astype(U64)
        return ((a[:, 0::2] << U64(3)) + (m[:, 0::2] << U64(10))
                + (a[:, 1::2] << U64(17)) + (m[:, 1::2] << U64(24)))

    def header_of(self, col):
        """3 header bits contributed to the column on the right."""
        if self.name == "HGR":
            return (((col & U64(1 << 11)) >> U64(9))
                    ^ ((col & U64(3 << 17)) >> U64(17)))
        return (col & U64(7 << 28)) >> U64(28)

    def footer_of(self, col):
        """3 footer bits contributed to the column on the left."""
        if self.name == "HGR":
            return (((col & U64(1 << 10)) >> U64(10))
                    ^ ((col & U64(3 << 3)) >> U64(2))) << U64(19)
        return (col & U64(7 << 3)) << U64(28)

    def pack(self, main, aux=None):
        body = self.body(main, aux)
        header = self.header_of(np.roll(body, 1, axis=1))
        header[:, 0] = 0
        footer = self.footer_of(np.roll(body, -1, axis=1))
        footer[:, -1] = 0
        return header ^ body ^ footer

    # -- per-byte views ----------------------------------------------------
    def byte_offset(self, page_offset: int, is_aux: bool) -> int:
        odd = page_offset & 1
        if self.name == "HGR":
            assert not is_aux
            return odd
        return (2 if odd else 0) if is_aux else (3 if odd else 1)

    def byte_offsets(self, is_aux: bool):
        if self.name == "HGR":
            assert not is_aux
            return (0, 1)
        return (0, 2) if is_aux else (1, 3)

    def mask_shift(self, data, o: int):
        return (data & self.masks[o]) >> self.shifts[o]

    def masked_update(self, o: int, old, value):
        if self.name == "HGR":
            if o == 0:
                return (old & ~U64(0xFF << 3)) ^ (U64(value) << U64(3))
            v = int(value)
            rot = ((v & 0x7F) << 1) ^ ((v & 0x80) >> 7)
            return (old & ~U64(0xFF << 11)) ^ (U64(rot) << U64(11))
        sh = 7 * o + 3
        return (old & ~U64(0x7F << sh)) ^ ((U64(value) & U64(0x7F)) << U64(sh))


SPECS = {"HGR": ModeSpec("HGR"), "DHGR": ModeSpec("DHGR")}


class OracleBitmap:
    """Bitmap / HGRBitmap / DHGRBitmap restated.  ``table`` is the symmetric
    uint16[(n_offsets, 4**bits)] array Bitmap.edit_distances returns."""

    def __init__(self, mode: str, table, main, aux=None):
        self.spec = SPECS[mode]
        self.table = table
        self.main = main          # uint8[32,256], aliased like MemoryMap
        self.aux = aux
        self.packed = self.spec.pack(main, aux)

    def repack(self):
        self.packed = self.spec.pack(self.main, self.aux)

    def apply(self, page, offset, is_aux, value):
        s = self.spec
        o = s.byte_offset(offset, is_aux)
        c = offset // 2
        self.packed[page, c] = s.masked_update(o, self.packed[page, c], value)
        if o == 0 and c > 0:
            self.packed[page, c - 1] = (
                (self.packed[page, c - 1] & s.keep_low)
                ^ s.footer_of(self.packed[page, c]))
        elif o == s.n_offsets - 1 and c < 127:
            self.packed[page, c + 1] = (
                (self.packed[page, c + 1] & s.keep_high)
                ^ s.header_of(self.packed[page, c]))
        (self.aux if is_aux else self.main)[page, offset] = value

    def _fix_array_neighbours(self, ary, o):
        s = self.spec
        if o == 0:
            ary &= s.keep_low
            ary ^= s.footer_of(np.roll(ary, -1, axis=1))
        elif o == s.n_offsets - 1:
            ary &= s.keep_high
            ary ^= s.header_of(np.roll(ary, 1, axis=1))

    def _diff(self, source_packed, target_packed, is_aux, content, shape):
        s = self.spec
        diff = np.ndarray(shape, dtype=np.int32)
        parts = []
        for o in s.byte_offsets(is_aux):
            if content is not None:
                cmp_packed = s.masked_update(o, source_packed, content)
                self._fix_array_neighbours(cmp_packed, o)
            else:
                cmp_packed = source_packed
            src = s.mask_shift(cmp_packed, o)
            tgt = s.mask_shift(target_packed, o)
            pair = (src << U64(s.masked_bits)) + tgt
            parts.append(self.table[o][pair].reshape(pair.shape))
        diff[..., 0::2] = parts[0]
        diff[..., 1::2] = parts[1]
        return diff

    def diff_weights(self, source: "OracleBitmap", is_aux, content=None):
        return self._diff(source.packed, self.packed, is_aux, content, (32, 256))

    def diff_weights_page(self, source_packed, target_packed, is_aux,
                          content=None):
        return self._diff(source_packed.reshape(1, -1),
                          target_packed.reshape(1, -1), is_aux, content,
                          (1, 256)).reshape(256)

    def compute_delta_page(self, page, content, diff_row, is_aux):
        row = self.packed[page, :]
        return self.diff_weights_page(row, row, is_aux, content) - diff_row

    def byte_pair_difference(self, o, old_packed, content):
        s = self.spec
        old = s.mask_shift(U64(old_packed), o)
        new = s.mask_shift(s.masked_update(o, U64(old_packed), content), o)
        return self.table[o][(old << U64(s.masked_bits)) + new]


class OracleVideo:
    """video.Video restated: greedy prioritised delta coder.

    RNG: by default the process-global ``random`` and ``np.random`` streams,
    exactly like the reference (video.py:178, 265, 291); tests may inject
    private generators with the same interfaces.
    """

    def __init__(self, mode: str, table, py_rng=None, np_rng=None):
        self.mode = mode
        self.spec = SPECS[mode]
        self.table = table
        self.main = np.zeros((32, 256), dtype=np.uint8)
        self.aux = np.zeros((32, 256), dtype=np.uint8) if mode == "DHGR" else None
        self.pixelmap = OracleBitmap(mode, table, self.main, self.aux)
        self.update_priority = np.zeros((32, 256), dtype=np.int32)
        self.aux_update_priority = (
            np.zeros((32, 256), dtype=np.int32) if mode == "DHGR" else None)
        self.out_of_work = {True: False, False: False}
        self._py = py_rng if py_rng is not None else _pyrandom
        self._np = np_rng if np_rng is not None else np.random
        self.mean_priority = 0.0

    def target_bitmap(self, main, aux=None) -> OracleBitmap:
        return OracleBitmap(self.mode, self.table, main, aux)

    def encode_frame(self, target: OracleBitmap, is_aux: bool):
        mem = self.aux if is_aux else self.main
        prio = self.aux_update_priority if is_aux else self.update_priority
        assert np.count_nonzero(mem[SCREEN_HOLES]) == 0
        self.mean_priority = float(prio.mean())   # video.py:90 prints this
        yield from self._index_changes(target, prio, is_aux)

    def _heapify(self, prio):
        pages, offsets = prio.nonzero()
        nonces = self._np.randint(0, 256, size=pages.shape[0])
        heap = [tuple(r) for r in np.stack(
            (-prio[pages, offsets], nonces, pages, offsets)).T.tolist()]
        heapq.heapify(heap)
        return heap

    def _candidates(self, page, content, target, diff, is_aux):
        delta = target.compute_delta_page(page, content, diff[page, :], is_aux)
        neg = delta < 0
        offs = np.arange(256)[neg]
        vals = delta[neg]
        heap = [(vals[i], self._py.getrandbits(8), offs[i])
                for i in range(len(offs))]
        heapq.heapify(heap)
        while heap:
            d, _, o = heapq.heappop(heap)
            yield -d, o

    def _index_changes(self, target, prio, is_aux):
        tmem = target.aux if (self.mode == "DHGR" and is_aux) else target.main
        diff = target.diff_weights(self.pixelmap, is_aux)
        diff[SCREEN_HOLES] = 0
        prio[diff == 0] = 0
        prio += diff
        assert np.all(prio >= 0)
        heap = self._heapify(prio)
        while heap:
            _, _, page, offset = heapq.heappop(heap)
            assert not SCREEN_HOLES[page, offset]
            if prio[page, offset] == 0:
                continue
            offsets = [offset]
            content = tmem[page, offset]
            if self.mode == "DHGR":
                assert content < 0x80
            prio[page, offset] = 0
            diff[page, offset] = 0
            self.pixelmap.apply(page, offset, is_aux, content)
            for _, o in self._candidates(page, content, target, diff, is_aux):
                assert o != offset and not SCREEN_HOLES[page, o]
                if prio[page, o] == 0:
                    continue
                bo = self.spec.byte_offset(int(o), is_aux)
                p = target.byte_pair_difference(
                    bo, target.packed[page, o // 2], content)
                prio[page, o] = p
                self.pixelmap.apply(page, o, is_aux, content)
                if p:
                    # -p on np.uint16 wraps to 65536-p (video.py:178): re-queued
                    # cells sort behind every first-pass cell.
                    heapq.heappush(heap, (
                        (65536 - int(p)) & 0xFFFF, self._py.getrandbits(8),
                        page, o))
                offsets.append(o)
                if len(offsets) == 3:
                    break
            while len(offsets) < 4:
                offsets.append(offsets[0])
            yield page + 32, content, offsets
        self.out_of_work[is_aux] = True
        content = tmem[0, 0]
        while True:
            yield 32, content, [0, 0, 0, 0]

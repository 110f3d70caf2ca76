"""Edit-distance table generation (reference transcoder/make_data_tables.py) on the
GPU, behind the reference's function names and signatures.

    python -m iivision_b200.make_data_tables        # writes transcoder/data/*.npz

colormath and weighted-levenshtein are not dependencies: the CIE2000 matrix is an
FP64 device kernel (iiv_lut_cie2000) and the weighted Damerau-Levenshtein
distances are the table kernels of csrc/iiv_tables.cu.  The reference's ~90
CPU-minutes (README.md:64-67) become milliseconds of kernel time; what remains is
the device->host copy and the compressed .npz write, whose deflate runs on all
host cores (npz_io.py).
"""

import functools
import os
import zlib
from typing import Iterable, Type

import numpy as np
import torch

from . import colours
from . import npz_io
from . import ops
from . import palette
from . import screen

PIXEL_CHARS = "0123456789ABCDEF"
DATA_DIR = "transcoder/data"


def pixel_char(i: int) -> str:
    return PIXEL_CHARS[i]


@functools.lru_cache(None)
def pixel_string(pixels: Iterable[int]) -> str:
    return "".join(pixel_char(p) for p in pixels)


class EditDistanceParams:
    """Data class for parameters to Damerau-Levenshtein edit distance.

    As in the reference (make_data_tables.py:30-52) the arrays are CLASS
    attributes that compute_substitute_costs fills in place.
    """

    # Insertions and deletions make no sense for pixel strings
    insert_costs = np.ones(128, dtype=np.float64) * 100000
    delete_costs = np.ones(128, dtype=np.float64) * 100000
    transpose_costs = np.ones((128, 128), dtype=np.float64)
    substitute_costs = np.zeros((128, 128), dtype=np.float64)
    # 5x costs for evaluating other offsets for a content byte (unused upstream)
    error_substitute_costs = np.zeros((128, 128), dtype=np.float64)


def compute_diff_matrix(pal: Type[palette.BasePalette]):
    """Matrix of int()-truncated CIE2000 delta-E between palette colour pairs,
    indexed by nominal colour value (device FP64; make_data_tables.py:55-70)."""
    return ops.lut_cie2000(pal.rgb_by_value())


def compute_substitute_costs(pal: Type[palette.BasePalette]):
    """Compute costs for substituting one colour pixel for another."""
    edp = EditDistanceParams()
    diff_matrix = compute_diff_matrix(pal)
    idx = np.frombuffer(PIXEL_CHARS.encode("ascii"), dtype=np.uint8)
    # (j, i) is written after (i, j) upstream, so S[c][d] = dm[max, min]; the
    # matrix is symmetric after truncation so the order is immaterial
    cost = np.tril(diff_matrix) + np.tril(diff_matrix, -1).T
    edp.substitute_costs[np.ix_(idx, idx)] = cost
    edp.error_substitute_costs[np.ix_(idx, idx)] = 5 * cost
    return edp


def _lut16(costs: np.ndarray) -> np.ndarray:
    idx = np.frombuffer(PIXEL_CHARS.encode("ascii"), dtype=np.uint8)
    sub = np.asarray(costs)[np.ix_(idx, idx)]
    lut = sub.astype(np.int32)
    if not np.array_equal(lut, sub):
        raise ValueError("substitution costs must be integral")
    return lut


def _check_chain_conditions(edp, n: int, lut: np.ndarray) -> None:
    """The device kernels evaluate the substitution/transposition chain the full
    Damerau-Levenshtein DP collapses to when an insert+delete pair can never pay
    for itself (make_data_tables.py:35-36).  Refuse parameter sets where that does
    not hold instead of silently computing something else."""
    idx = np.frombuffer(PIXEL_CHARS.encode("ascii"), dtype=np.uint8)
    if (min(edp.insert_costs[idx].min(), edp.delete_costs[idx].min())
            <= n * max(int(lut.max()), 1)):
        raise ValueError("insert/delete costs too small for the substitution-only "
                         "collapse; unsupported parameter set")
    tc = getattr(edp, "transpose_costs", None)
    if tc is not None and not np.all(np.asarray(tc)[np.ix_(idx, idx)] == 1):
        raise ValueError("only unit transposition costs are supported")
    if not np.array_equal(lut, lut.T) or np.diagonal(lut).any():
        raise ValueError("substitution costs must be symmetric with zero diagonal")


def edit_distance(edp: EditDistanceParams, a: str, b: str, error: bool) -> np.float64:
    """Damerau-Levenshtein edit distance between two pixel strings."""
    if len(a) != len(b):
        raise ValueError("pixel strings must have equal length")
    lut = _lut16(edp.error_substitute_costs if error else edp.substitute_costs)
    _check_chain_conditions(edp, max(len(a), 1), lut)
    sa = np.array([[PIXEL_CHARS.index(c) for c in a]], dtype=np.uint8).reshape(1, -1)
    sb = np.array([[PIXEL_CHARS.index(c) for c in b]], dtype=np.uint8).reshape(1, -1)
    res = np.float64(ops.string_distance(lut, sa, sb)[0]) if len(a) else np.float64(0)
    assert (0 <= res < 2 ** 16), res
    return res


def _mode_of(bitmap_cls) -> int:
    name = getattr(bitmap_cls, "NAME", None)
    if name not in ops.MODES:
        raise ValueError("unknown bitmap class %r" % (bitmap_cls,))
    m = ops.MODES[name]
    if (int(bitmap_cls.MASKED_BITS) != ops.MASKED_BITS[m]
            or int(bitmap_cls.MASKED_DOTS) != ops.MASKED_DOTS[m]
            or list(bitmap_cls.PHASES) != ops.mode_phases(m)):
        raise ValueError("%s geometry differs from the compiled kernels" % name)
    return m


def compute_edit_distance_device(edp: EditDistanceParams, bitmap_cls,
                                 layout: int = ops.LAYOUT_TRIANGULAR,
                                 out: torch.Tensor = None) -> torch.Tensor:
    """compute_edit_distance leaving the table in HBM (uint16[n_off, 4**bits])."""
    m = _mode_of(bitmap_cls)
    lut = _lut16(edp.substitute_costs)
    _check_chain_conditions(edp, ops.MASKED_DOTS[m], lut)
    return ops.table_generate(m, lut, layout=layout, out=out)


def compute_edit_distance(edp: EditDistanceParams, bitmap_cls: Type[screen.Bitmap],
                          nominal_colours: Type[colours.NominalColours] = None
                          ) -> np.ndarray:
    """Computes edit distance matrix between all pairs of pixel strings.

    Returns the reference's array: uint16[(len(BYTE_MASKS), 4**MASKED_BITS)],
    entry (i << bits) + j filled for j < i only (make_data_tables.py:156-172).
    ``nominal_colours`` only validated pixel values upstream and is ignored.

    The array is READ-ONLY (``np.array(result)`` gives a private writable copy): it is a
    view of recycled page-locked memory whose upper triangle is known to be zero, which is
    what lets the device-to-host copy skip it.  The reference's callers only read the
    result (make_data_tables.py:186-188 saves it).
    """
    m = _mode_of(bitmap_cls)
    table = compute_edit_distance_device(edp, bitmap_cls, ops.LAYOUT_TRIANGULAR)
    host, arr = _pinned_pool.lease(tuple(table.shape))
    ops.table_download(m, table, host.data_ptr(), layout=ops.LAYOUT_TRIANGULAR)
    torch.cuda.current_stream().synchronize()
    return arr


class _PinnedPool:
    """Page-locked uint16 buffers for tables on their way to the host.  cudaHostAlloc of
    1 GiB costs tens of milliseconds, more than the copy itself, so buffers are recycled: a
    lease hands out a read-only array over a buffer and the buffer returns to the pool when
    that array -- and with it every view derived from it, which numpy chains to it through
    ``.base`` -- has been garbage collected.  Buffers start zeroed and are only ever written
    below the diagonal, so a recycled buffer is still zero everywhere else."""

    def __init__(self, keep: int = 4):
        self.keep = keep
        self.free = []

    def lease(self, shape):
        import weakref
        host = None
        for k, t in enumerate(self.free):
            if tuple(t.shape) == shape:
                host = self.free.pop(k)
                break
        if host is None:
            host = self._alloc(shape)
        arr = host.numpy()
        arr.flags.writeable = False
        weakref.finalize(arr, self._release, host)
        return host, arr

    @staticmethod
    def _alloc(shape):
        return torch.zeros(shape, dtype=torch.uint16, pin_memory=True)

    def _release(self, host):
        if len(self.free) < self.keep:
            self.free.append(host)


_pinned_pool = _PinnedPool()


DEFLATE_GROUP = 64      # device blocks (32 KiB each) per piece of load_member's index
STAGE_SLOTS, STAGE_BYTES = 3, 16 << 20


class _StageRings:
    """Rings of page-locked staging slots, one ring per stream on its way home at a time
    (make_data_tables.main writes several files at once); allocating a ring costs ~10 ms of
    cudaHostAlloc, so finished rings are kept for the next stream."""

    def __init__(self):
        import threading
        self.lock = threading.Lock()
        self.free = {}

    def take(self, dev):
        with self.lock:
            rings = self.free.setdefault(dev, [])
            if rings:
                return rings.pop()
        return [torch.empty(STAGE_BYTES, dtype=torch.uint8, pin_memory=True)
                for _ in range(STAGE_SLOTS)]

    def give(self, dev, ring):
        with self.lock:
            self.free.setdefault(dev, []).append(ring)


_rings = _StageRings()


def _staged_home(stream: torch.Tensor):
    """Yields a device byte stream as host buffers, in order, through a small ring of
    page-locked slots: while the caller consumes one slice (writes it to the file) the next
    ones are already crossing PCIe.  A yielded buffer is valid until the next is asked for.
    (Page-locking memory for the whole stream would cost more than the copy.)"""
    dev = stream.device
    slots = _rings.take(dev)
    try:
        total = stream.numel()
        n = (total + STAGE_BYTES - 1) // STAGE_BYTES
        events = [None] * n

        def issue(k):
            lo = k * STAGE_BYTES
            hi = min(total, lo + STAGE_BYTES)
            slots[k % STAGE_SLOTS][:hi - lo].copy_(stream[lo:hi], non_blocking=True)
            events[k] = torch.cuda.Event()
            events[k].record()

        for k in range(min(n, STAGE_SLOTS)):
            issue(k)
        for k in range(n):
            events[k].synchronize()
            hi = min(total, (k + 1) * STAGE_BYTES)
            yield slots[k % STAGE_SLOTS].numpy()[:hi - k * STAGE_BYTES]
            if k + STAGE_SLOTS < n:
                issue(k + STAGE_SLOTS)       # the consumer is done with slot k
    finally:
        torch.cuda.current_stream().synchronize()
        _rings.give(dev, slots)


def make_edit_distance(pal: Type[palette.BasePalette], edp: EditDistanceParams,
                       bitmap_cls: Type[screen.Bitmap],
                       nominal_colours: Type[colours.NominalColours] = None,
                       _writer=None):
    """Write file containing (D)HGR edit distance matrix for a palette.

    Same container and member as the reference's ``np.savez_compressed(data,
    edit_distance=dist)`` (make_data_tables.py:186-188) -- ``np.load(...)['edit_distance']``
    reads it -- but the table never comes to the host uncompressed: the device deflates it
    where it was generated (ops.deflate_table) and only the compressed stream crosses PCIe,
    slice by slice, while the slices before it are being written to the file.

    ``_writer`` (main()'s): called with the function that brings the stream home and writes
    the file, instead of that function being run here -- the next table is then generated
    and deflated while this one is still being written."""
    from . import deflate
    m = _mode_of(bitmap_cls)
    table = compute_edit_distance_device(edp, bitmap_cls, ops.LAYOUT_TRIANGULAR)
    stream, sizes, block_crc, block_bytes = ops.deflate_table(m, table)
    shape = tuple(table.shape)
    del table                       # deflate_table has synchronised: the stream is complete
    data = "%s/%s_palette_%d_edit_distance.npz" % (
        DATA_DIR, bitmap_cls.NAME, pal.ID.value)

    def write_file():
        home = _staged_home(stream)
        first = next(home)          # starts the copies; the bookkeeping below overlaps them
        # checksums and the piece index
        like = np.broadcast_to(np.uint16(0), shape)
        header = npz_io._npy_header(like)
        body_crc = deflate.crc32_of_equal_parts(block_crc, block_bytes)
        crc = deflate.crc32_combine(zlib.crc32(header), body_crc, len(block_crc) * block_bytes)
        group = DEFLATE_GROUP if len(sizes) % DEFLATE_GROUP == 0 else 1
        ends = np.cumsum(sizes.astype(np.int64))
        comp_len = ends[group - 1::group] - np.concatenate(([0], ends[group - 1::group][:-1]))
        comp_off = ends[group - 1::group] - comp_len
        raw_len = group * block_bytes
        group_crc = deflate.crc32_of_groups(block_crc, block_bytes, group)
        pieces = [(k * raw_len, raw_len, int(comp_off[k]), int(comp_len[k]), int(group_crc[k]))
                  for k in range(len(comp_len))]
        import itertools
        try:
            npz_io.savez_predeflated(data, "edit_distance", like, itertools.chain((first,), home),
                                     stream.numel(), crc, pieces)
        finally:
            home.close()            # the staging ring goes back whatever happened

    if _writer is None:
        write_file()
    else:
        _writer(write_file)


def table_jobs():
    """The files make_data_tables.main() writes (make_data_tables.py:191-204), in its order:
    (palette, bitmap class, nominal colours, relative cost = table bytes)."""
    jobs = []
    for p in palette.PALETTES.values():
        jobs.append((p, screen.HGRBitmap, colours.HGRColours, 2.0))
        jobs.append((p, screen.DHGRBitmap, colours.DHGRColours, 1.0))
    return jobs


def main(rank: int = None, world: int = None):
    """All table files.  Under torchrun (RANK / WORLD_SIZE set, or given) the files are
    independent jobs and are shared out over the ranks, one GPU each; nothing is exchanged."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    os.makedirs(DATA_DIR, mode=0o755, exist_ok=True)
    jobs = table_jobs()
    mine = range(len(jobs))
    if world > 1:
        from . import parallel
        mine = parallel.shard_jobs([j[3] for j in jobs], world, rank)
    edps = {}
    written = []
    # A file's way home (PCIe slices + the host's file writes, ~0.1 s) runs on a thread of
    # its own, with its own CUDA stream, while this thread generates and deflates the next
    # table (~25 ms of device time): the files of a rank are written side by side.
    import threading
    workers, failures = [], []

    def in_background(write_file):
        device = torch.cuda.current_device()

        def run():
            try:
                torch.cuda.set_device(device)
                with torch.cuda.stream(torch.cuda.Stream()):
                    write_file()
            except BaseException as e:   # noqa: BLE001 -- re-raised by the caller's thread
                failures.append(e)

        t = threading.Thread(target=run, name="iiv-table-writer")
        t.start()
        workers.append(t)

    try:
        for k in mine:
            p, bitmap_cls, nominal, _ = jobs[k]
            if p not in edps:
                print("Processing palette %s" % p)
                edps[p] = compute_substitute_costs(p)
            make_edit_distance(p, edps[p], bitmap_cls, nominal, _writer=in_background)
            written.append("%s/%s_palette_%d_edit_distance.npz"
                           % (DATA_DIR, bitmap_cls.NAME, p.ID.value))
    finally:
        for t in workers:
            t.join()
    if failures:
        raise failures[0]
    return written


if __name__ == "__main__":
    main()

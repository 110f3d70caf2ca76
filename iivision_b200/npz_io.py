"""``np.savez_compressed`` with the deflate work spread over the host's cores.

The reference stores each edit-distance table as a compressed ``.npz``
(make_data_tables.py:186-188) and loads it back with ``np.load`` (screen.py:352).
Once the table itself takes a quarter of a millisecond, zlib on one core (tens of
seconds per GiB) is the whole cost of ``make_edit_distance``.  A deflate stream may be
cut anywhere into independently compressed pieces as long as every piece but the last
ends in a sync flush (byte-aligned, not final) -- the pigz construction -- so the pieces
are compressed in a thread pool (zlib releases the GIL) and written in order into an
ordinary ZIP container.  ``np.load`` and any unzip read the result like numpy's own.

The writer also stores where its pieces start (member ``__deflate_pieces__``, which
readers that ask for ``edit_distance`` never see), so that ``load_member`` can inflate
them on all cores as well; files from ``np.savez_compressed`` itself are read through
``np.load``.
"""

import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

CHUNK = 8 << 20
_LOCAL = b"PK\x03\x04"
_CENTRAL = b"PK\x01\x02"
_END = b"PK\x05\x06"
_END64 = b"PK\x06\x06"
_END64_LOC = b"PK\x06\x07"
_DOS_DATE = (2019 - 1980) << 9 | 1 << 5 | 1     # fixed timestamp: files are reproducible


def _npy_header(a: np.ndarray) -> bytes:
    import io
    buf = io.BytesIO()
    np.lib.format.write_array_header_1_0(buf, np.lib.format.header_data_from_array_1_0(a))
    return buf.getvalue()


def _deflate_piece(args):
    view, level, last = args
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    out = co.compress(view)
    out += co.flush(zlib.Z_FINISH if last else zlib.Z_SYNC_FLUSH)
    return out


def _member(f, pool, name: str, a: np.ndarray, level: int):
    """Writes one .npy member; returns its central-directory record fields."""
    a = np.asarray(a, order="C")
    if a.dtype.hasobject:
        raise ValueError("object arrays are not supported")
    header = _npy_header(a)
    body = memoryview(a.reshape(-1).view(np.uint8)) if a.size else memoryview(b"")
    pieces = [memoryview(header)]
    pieces += [body[o:o + CHUNK] for o in range(0, len(body), CHUNK)]
    jobs = [(p, level, k == len(pieces) - 1) for k, p in enumerate(pieces)]
    raw_size = len(header) + len(body)
    fname = (name + ".npy").encode("utf-8")
    offset = f.tell()
    # sizes and CRC are not known yet: ZIP64 extra field reserved, patched afterwards
    extra = struct.pack("<HHQQ", 1, 16, 0, 0)
    f.write(_LOCAL + struct.pack("<HHHHHIIIHH", 45, 0, 8, 0, _DOS_DATE, 0, 0xffffffff,
                                 0xffffffff, len(fname), len(extra)) + fname + extra)
    crc = 0
    comp_size = 0
    crc_done = 0
    index = []          # (raw offset, raw length, file offset, compressed length, crc32)
    for k, out in enumerate(pool.map(_deflate_piece, jobs)):
        index.append((crc_done, len(pieces[k]), f.tell(), len(out), zlib.crc32(pieces[k])))
        f.write(out)
        comp_size += len(out)
        # the member's checksum runs over the raw bytes, piece by piece, while later
        # pieces are still being compressed by the pool
        crc = zlib.crc32(pieces[k], crc)
        crc_done += len(pieces[k])
    assert crc_done == raw_size
    end = f.tell()
    f.seek(offset + 14)
    f.write(struct.pack("<I", crc))
    f.seek(offset + 30 + len(fname) + 4)
    f.write(struct.pack("<QQ", raw_size, comp_size))
    f.seek(end)
    return (fname, crc, raw_size, comp_size, offset), index


PIECES = "__deflate_pieces__"


def savez_predeflated(path, name: str, array_like, stream, stream_len: int, raw_crc: int,
                      pieces, level: int = 6) -> None:
    """An ``.npz`` with ONE member whose body has already been deflated elsewhere (on the
    device: ops.deflate_table).

    ``array_like`` only describes the array (shape, dtype, C order: anything
    ``header_data_from_array_1_0`` accepts, e.g. an ``np.broadcast_to`` view); ``stream``
    yields the body's raw-deflate bytes in order, ``stream_len`` bytes in all, as buffers that
    are only valid until the next one is asked for (a staging ring) -- a sequence of
    byte-aligned, self-contained non-final blocks; ``raw_crc`` is the CRC-32 of the .npy
    header + body, i.e. of the member; ``pieces`` rows of (raw offset in the body, raw
    length, offset in the stream, compressed length, CRC-32 of those raw bytes) for
    ``load_member``'s parallel inflate."""
    path = os.fspath(path)
    if not path.endswith(".npz"):
        path += ".npz"
    header = _npy_header(array_like)
    head = _deflate_piece((memoryview(header), level, False))
    tail = b"\x01\x00\x00\xff\xff"            # final, empty stored block
    body_raw = int(np.prod(array_like.shape)) * array_like.dtype.itemsize
    raw_size = len(header) + body_raw
    comp_size = len(head) + int(stream_len) + len(tail)
    fname = (name + ".npy").encode("utf-8")
    with open(path, "wb") as f:
        extra = struct.pack("<HHQQ", 1, 16, raw_size, comp_size)
        f.write(_LOCAL + struct.pack("<HHHHHIIIHH", 45, 0, 8, 0, _DOS_DATE, raw_crc,
                                     0xffffffff, 0xffffffff, len(fname), len(extra))
                + fname + extra)
        data_start = f.tell()
        f.write(head)
        body_start = f.tell()
        written = 0
        for part in stream:
            part = memoryview(part)
            f.write(part)
            written += len(part)
        if written != int(stream_len):
            raise ValueError("stream of %d bytes, %d announced" % (written, stream_len))
        f.write(tail)
        table = [(0, 0, len(header), data_start, len(head), zlib.crc32(header))]
        for raw_off, raw_len, comp_off, comp_len, crc in pieces:
            table.append((0, len(header) + int(raw_off), int(raw_len),
                          body_start + int(comp_off), int(comp_len), int(crc)))
        if len(table) > 1:
            # the final empty block rides with the last piece: the index covers the member
            last = table[-1]
            table[-1] = last[:4] + (last[4] + len(tail),) + last[5:]
        records = [(fname, raw_crc, raw_size, comp_size, 0)]
        with ThreadPoolExecutor(max_workers=1) as pool:
            rec, _ = _member(f, pool, PIECES, np.array(table, dtype=np.int64).reshape(-1, 6),
                             level)
        records.append(rec)
        cd_start = f.tell()
        for fn, crc, rs, cs, offset in records:
            extra = struct.pack("<HHQQQ", 1, 24, rs, cs, offset)
            f.write(_CENTRAL + struct.pack(
                "<HHHHHHIIIHHHHHII", 45, 45, 0, 8, 0, _DOS_DATE, crc, 0xffffffff, 0xffffffff,
                len(fn), len(extra), 0, 0, 0, 0, 0xffffffff) + fn + extra)
        cd_size = f.tell() - cd_start
        n = len(records)
        f.write(_END64 + struct.pack("<QHHIIQQQQ", 44, 45, 45, 0, 0, n, n, cd_size, cd_start))
        f.write(_END64_LOC + struct.pack("<IQI", 0, cd_start + cd_size, 1))
        f.write(_END + struct.pack("<HHHHIIH", 0, 0, min(n, 0xffff), min(n, 0xffff),
                                   0xffffffff, 0xffffffff, 0))


def savez_compressed(path, level: int = 6, threads: int = None, index: bool = True,
                     **arrays) -> None:
    """Drop-in for ``np.savez_compressed(path, **arrays)`` (keyword form).  With
    ``index`` an extra int64[n, 6] member lists (member number, raw offset, raw length,
    file offset, compressed length, CRC-32 of the raw bytes) of every deflate piece for
    ``load_member``."""
    path = os.fspath(path)
    if not path.endswith(".npz"):
        path += ".npz"
    threads = threads or min(32, os.cpu_count() or 1)
    records = []
    with open(path, "wb") as f, ThreadPoolExecutor(max_workers=threads) as pool:
        if PIECES in arrays:
            raise ValueError("%s is a reserved member name" % PIECES)
        table = []
        for k, (name, a) in enumerate(arrays.items()):
            rec, pieces = _member(f, pool, name, np.asanyarray(a), level)
            records.append(rec)
            table += [(k,) + p for p in pieces]
        if index:
            rec, _ = _member(f, pool, PIECES, np.array(table, dtype=np.int64).reshape(-1, 6),
                             level)
            records.append(rec)
        cd_start = f.tell()
        for fname, crc, raw_size, comp_size, offset in records:
            extra = struct.pack("<HHQQQ", 1, 24, raw_size, comp_size, offset)
            f.write(_CENTRAL + struct.pack(
                "<HHHHHHIIIHHHHHII", 45, 45, 0, 8, 0, _DOS_DATE, crc, 0xffffffff, 0xffffffff,
                len(fname), len(extra), 0, 0, 0, 0, 0xffffffff) + fname + extra)
        cd_size = f.tell() - cd_start
        n = len(records)
        f.write(_END64 + struct.pack("<QHHIIQQQQ", 44, 45, 45, 0, 0, n, n, cd_size, cd_start))
        f.write(_END64_LOC + struct.pack("<IQI", 0, cd_start + cd_size, 1))
        f.write(_END + struct.pack("<HHHHIIH", 0, 0, min(n, 0xffff), min(n, 0xffff),
                                   0xffffffff, 0xffffffff, 0))


def load_member(path, name: str, threads: int = None, alloc=None) -> np.ndarray:
    """``np.load(path)[name]``.  Files written by ``savez_compressed`` above are inflated
    piece by piece on ``threads`` cores, into the buffer ``alloc(n_bytes)`` returns (a
    writable uint8 array, e.g. page-locked memory) when given; any other file goes through
    ``np.load``."""
    import io
    import mmap
    import zipfile
    path = os.fspath(path)
    with zipfile.ZipFile(path) as z:
        names = z.namelist()
        if name + ".npy" not in names:
            raise KeyError("%s is not a file in the archive" % name)
        if PIECES + ".npy" not in names:
            with np.load(path) as npz:
                return npz[name]
        member = names.index(name + ".npy")
        info = z.getinfo(name + ".npy")
        with z.open(PIECES + ".npy") as fp:
            table = np.lib.format.read_array(fp)
    rows = table[table[:, 0] == member] if table.ndim == 2 and table.shape[1] == 6 else table[:0]
    raw_size = int(info.file_size)
    if (table.ndim != 2 or table.shape[1] != 6 or rows.size == 0
            or int(rows[:, 2].sum()) != raw_size
            or int(rows[:, 4].sum()) != int(info.compress_size)):
        with np.load(path) as npz:       # index does not describe this member: plain path
            return npz[name]
    out = alloc(raw_size) if alloc is not None else np.empty(raw_size, dtype=np.uint8)
    threads = threads or min(32, os.cpu_count() or 1)
    with open(path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as mm:
        view = memoryview(mm)

        def inflate(row):
            _, raw_off, raw_len, file_off, comp_len, crc = (int(x) for x in row)
            data = zlib.decompressobj(-15).decompress(view[file_off:file_off + comp_len])
            if len(data) != raw_len or zlib.crc32(data) != crc:
                raise ValueError("corrupt deflate piece at file offset %d" % file_off)
            out[raw_off:raw_off + raw_len] = np.frombuffer(data, dtype=np.uint8)

        try:
            with ThreadPoolExecutor(max_workers=threads) as pool:
                list(pool.map(inflate, rows))
        finally:
            view.release()
    head = io.BytesIO(out[:int(rows[0, 2])].tobytes())
    version = np.lib.format.read_magic(head)
    if version != (1, 0):
        raise ValueError("unexpected .npy version %r" % (version,))
    shape, fortran, dtype = np.lib.format.read_array_header_1_0(head)
    body = out[head.tell():]
    a = body.view(dtype)
    return a.reshape(shape, order="F" if fortran else "C")

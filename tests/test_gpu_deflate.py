"""GPU: the device-side deflate of a table (csrc/iiv_deflate.cu, next row N1) round-trips
through zlib -- the reference's loader is np.load (screen.py:352) -- and its CRCs are right."""

import os
import zipfile
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from iivision_b200 import ops
    return ops


@pytest.mark.parametrize("mode,pid,layout", [("DHGR", 5, 0), ("HGR", 5, 0), ("DHGR", 0, 1)])
def test_deflate_round_trip(ops, oracle_luts, mode, pid, layout):
    import torch
    from iivision_b200 import deflate
    table = ops.table_generate(mode, oracle_luts[pid], layout=layout)
    stream, sizes, crcs, block_bytes = ops.deflate_table(mode, table)
    torch.cuda.synchronize()
    raw = table.cpu().view(torch.int16).numpy().tobytes()
    comp = stream.cpu().numpy().tobytes()
    assert int(sizes.sum()) == len(comp) and len(sizes) * block_bytes == len(raw)
    got = zlib.decompressobj(-15).decompress(comp + b"\x01\x00\x00\xff\xff")
    assert got == raw
    # every block stands alone (the writer's piece index relies on it) and its CRC is right
    ends = np.cumsum(sizes.astype(np.int64))
    for b in (0, 1, len(sizes) // 2, len(sizes) - 1):
        piece = comp[int(ends[b] - sizes[b]):int(ends[b])]
        one = zlib.decompressobj(-15).decompress(piece)
        assert one == raw[b * block_bytes:(b + 1) * block_bytes]
        assert int(crcs[b]) == zlib.crc32(one)
    assert deflate.crc32_of_equal_parts(crcs, block_bytes) == zlib.crc32(raw)
    # size against zlib level 6 on 8 MiB from the middle rows of the first offset's slice
    first = len(sizes) // table.shape[0] // 2
    lo = first * block_bytes
    sample = raw[lo:lo + 256 * block_bytes]
    ours = int(sizes[first:first + 256].sum())
    theirs = len(zlib.compress(sample, 6))
    print("%s pid %d layout %d: ratio %.3f, zlib-6 %.3f on the mid-table sample, whole table %.3f"
          % (mode, pid, layout, len(sample) / ours, len(sample) / theirs, len(raw) / len(comp)))
    # the file holds the triangular layout (layout 0); the symmetric one has no zeros to win on
    assert ours < (1.3 if layout == 0 else 1.45) * theirs


def test_make_edit_distance_file_is_a_plain_npz(ops, oracle_luts, tmp_path, monkeypatch):
    from iivision_b200 import colours, make_data_tables as mdt, npz_io, palette, screen
    from oracle import tables
    monkeypatch.setattr(mdt, "DATA_DIR", str(tmp_path))
    pal = palette.NTSCPalette
    edp = mdt.compute_substitute_costs(pal)
    mdt.make_edit_distance(pal, edp, screen.HGRBitmap, colours.HGRColours)
    path = str(tmp_path / "HGR_palette_5_edit_distance.npz")
    want, _ = tables.build_table("HGR", oracle_luts[5], triangular=True)
    with np.load(path) as z:
        assert np.array_equal(z["edit_distance"], want)
    with zipfile.ZipFile(path) as z:
        assert z.testzip() is None
        assert z.getinfo("edit_distance.npy").file_size == want.nbytes + 128
    assert np.array_equal(npz_io.load_member(path, "edit_distance"), want)


def test_main_writes_the_four_files_side_by_side(oracle_luts, tmp_path, monkeypatch):
    """make_data_tables.main() (make_data_tables.py:191-204): every file is brought home and
    written on a thread of its own while the next table is generated and deflated; each must
    still be the oracle's table, and a second call over the same directory as well (staging
    rings recycled)."""
    import contextlib
    import io
    from iivision_b200 import make_data_tables as mdt, npz_io
    from oracle import tables
    monkeypatch.setattr(mdt, "DATA_DIR", str(tmp_path))
    for _ in range(2):
        with contextlib.redirect_stdout(io.StringIO()):
            written = mdt.main(0, 1)
        assert [os.path.basename(w) for w in written] == [
            "HGR_palette_0_edit_distance.npz", "DHGR_palette_0_edit_distance.npz",
            "HGR_palette_5_edit_distance.npz", "DHGR_palette_5_edit_distance.npz"]
    for w in written:
        mode, _, pid = os.path.basename(w).split("_")[:3]
        want, _ = tables.build_table(mode, oracle_luts[int(pid)], triangular=True)
        assert np.array_equal(npz_io.load_member(w, "edit_distance"), want), w
        with zipfile.ZipFile(w) as z:
            assert z.testzip() is None

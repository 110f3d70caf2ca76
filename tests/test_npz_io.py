"""Host logic: the multi-threaded .npz writer produces files np.load / zipfile read like
numpy's own (reference reader: screen.py:352, np.load(...)["edit_distance"])."""

import importlib.util
import os
import zipfile
import zlib

import numpy as np
import pytest


def _load_module():
    # iivision_b200/__init__ needs nothing from CUDA, but keep this test independent of it
    path = os.path.join(os.path.dirname(__file__), "..", "iivision_b200", "npz_io.py")
    spec = importlib.util.spec_from_file_location("npz_io_under_test", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("threads", [1, 4])
def test_round_trip_matches_numpy(tmp_path, threads):
    npz_io = _load_module()
    rng = np.random.default_rng(3)
    tri = np.tril(rng.integers(0, 2000, size=(3000, 3000), dtype=np.uint16))   # 18 MB: 3 pieces
    arrays = {
        "edit_distance": tri.reshape(2, -1),
        "empty": np.zeros((0, 5), np.int32),
        "scalar": np.float64(2.5),
        "strided": np.arange(100, dtype=np.int64)[::3],
        "fortran": np.asfortranarray(rng.integers(0, 9, size=(7, 5), dtype=np.int8)),
    }
    ours = tmp_path / "ours.npz"
    theirs = tmp_path / "theirs.npz"
    npz_io.savez_compressed(str(ours), threads=threads, **arrays)
    np.savez_compressed(str(theirs), **arrays)
    with np.load(str(ours)) as a, np.load(str(theirs)) as b:
        assert sorted(set(a.files) - {npz_io.PIECES}) == sorted(b.files)
        for k in b.files:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
            assert np.array_equal(a[k], b[k]), k
    with zipfile.ZipFile(str(ours)) as z:
        assert z.testzip() is None          # CRCs and sizes of every member check out
        info = z.getinfo("edit_distance.npy")
        assert info.compress_type == zipfile.ZIP_DEFLATED
        assert info.file_size == tri.nbytes + 128
    # within a few percent of numpy's single-stream deflate
    assert os.path.getsize(ours) < 1.05 * os.path.getsize(theirs) + 4096


def test_suffix_is_added_and_pieces_are_cut_at_chunk_size(tmp_path, monkeypatch):
    npz_io = _load_module()
    monkeypatch.setattr(npz_io, "CHUNK", 1000)
    data = np.arange(5000, dtype=np.uint16)
    npz_io.savez_compressed(str(tmp_path / "t"), edit_distance=data)
    with np.load(str(tmp_path / "t.npz")) as z:
        assert np.array_equal(z["edit_distance"], data)


def test_load_member_parallel_and_fallback(tmp_path, monkeypatch):
    npz_io = _load_module()
    monkeypatch.setattr(npz_io, "CHUNK", 4096)
    rng = np.random.default_rng(5)
    arrays = {"edit_distance": rng.integers(0, 300, size=(4, 9001), dtype=np.uint16),
              "f": np.asfortranarray(rng.integers(0, 9, size=(6, 5), dtype=np.int32)),
              "empty": np.zeros((0,), np.uint8)}
    ours, theirs = str(tmp_path / "ours.npz"), str(tmp_path / "theirs.npz")
    npz_io.savez_compressed(ours, **arrays)
    np.savez_compressed(theirs, **arrays)
    handed_out = []

    def alloc(n):
        handed_out.append(np.empty(n, np.uint8))
        return handed_out[-1]
    for path in (ours, theirs):
        for k, want in arrays.items():
            got = npz_io.load_member(path, k, threads=3, alloc=alloc)
            assert got.dtype == want.dtype and got.shape == want.shape, (path, k)
            assert np.array_equal(got, want), (path, k)
    assert len(handed_out) == 3           # the indexed file only; numpy's went through np.load
    with np.load(ours) as z:              # the reference's reader sees the member it asks for
        assert np.array_equal(z["edit_distance"], arrays["edit_distance"])
    with pytest.raises(KeyError):
        npz_io.load_member(ours, "missing")
    # a damaged piece is reported, not returned
    blob = bytearray(open(ours, "rb").read())
    with zipfile.ZipFile(ours) as z:
        start = z.getinfo("edit_distance.npy").header_offset
    blob[start + 400] ^= 0xff
    bad = str(tmp_path / "bad.npz")
    open(bad, "wb").write(bytes(blob))
    with pytest.raises((ValueError, zlib.error)):
        npz_io.load_member(bad, "edit_distance")


def test_reserved_member_name(tmp_path):
    npz_io = _load_module()
    with pytest.raises(ValueError):
        npz_io.savez_compressed(str(tmp_path / "x.npz"), **{npz_io.PIECES: np.zeros(1)})


def test_pinned_pool_recycles_only_dead_leases(monkeypatch):
    """make_data_tables._PinnedPool: a buffer returns to the pool when the leased array AND
    every view derived from it are gone (numpy chains views to the leased array)."""
    import gc
    torch = pytest.importorskip("torch")
    try:
        from iivision_b200 import make_data_tables as mdt
    except ImportError as e:       # library not built
        pytest.skip(str(e))
    pool = mdt._PinnedPool(keep=2)
    made = []

    def alloc(shape):
        made.append(1)
        return torch.zeros(shape, dtype=torch.uint16)
    monkeypatch.setattr(pool, "_alloc", alloc)
    host, a = pool.lease((2, 64))
    del host
    assert not a.flags.writeable and a.shape == (2, 64)
    with pytest.raises(ValueError):
        a[0, 0] = 1
    view = a[1, 3:9]
    ptr = a.ctypes.data
    del a
    gc.collect()
    _, b = pool.lease((2, 64))
    assert len(made) == 2 and b.ctypes.data != ptr        # the view still pins the first
    del view
    gc.collect()
    _, c = pool.lease((2, 64))
    assert len(made) == 2 and c.ctypes.data == ptr        # now recycled
    _, d = pool.lease((4, 8))
    assert len(made) == 3                                 # other shapes get their own


def test_predeflated_member_reads_like_numpys_and_by_pieces(tmp_path, monkeypatch):
    """savez_predeflated: the body arrives already deflated (on the GPU path: from the
    device) as byte-aligned non-final blocks, in slices; np.load reads the member and
    load_member inflates it by the piece index without falling back to np.load."""
    npz_io = _load_module()
    rng = np.random.default_rng(5)
    a = np.tril(rng.integers(0, 1800, size=(1024, 1024), dtype=np.uint16)).reshape(2, -1)
    raw = a.tobytes()
    part = 1 << 16
    blocks, pieces, off = [], [], 0
    for lo in range(0, len(raw), part):
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        out = co.compress(raw[lo:lo + part]) + co.flush(zlib.Z_SYNC_FLUSH)
        pieces.append((lo, part, off, len(out), zlib.crc32(raw[lo:lo + part])))
        blocks.append(out)
        off += len(out)
    stream = b"".join(blocks)
    slices = (stream[k:k + 100000] for k in range(0, len(stream), 100000))
    like = np.broadcast_to(np.uint16(0), a.shape)
    crc = zlib.crc32(npz_io._npy_header(like) + raw)
    path = str(tmp_path / "pre.npz")
    npz_io.savez_predeflated(path, "edit_distance", like, slices, len(stream), crc, pieces)
    with np.load(path) as z:
        assert np.array_equal(z["edit_distance"], a)
    with zipfile.ZipFile(path) as z:
        assert z.testzip() is None
    def no_fallback(*args, **kw):
        raise AssertionError("load_member fell back to np.load")
    monkeypatch.setattr(npz_io.np, "load", no_fallback)
    assert np.array_equal(npz_io.load_member(path, "edit_distance", threads=3), a)
    with pytest.raises(ValueError):
        npz_io.savez_predeflated(str(tmp_path / "short.npz"), "edit_distance", like,
                                 [stream[:-1]], len(stream), crc, pieces)

"""Host logic: the multi-threaded .npz writer produces files np.load / zipfile read like
numpy's own (reference reader: screen.py:352, np.load(...)["edit_distance"])."""

import importlib.util
import os
import zipfile

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
        assert sorted(a.files) == sorted(b.files)
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

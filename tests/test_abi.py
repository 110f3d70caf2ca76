"""CPU: the C-ABI library loads without a GPU and exports every symbol that
include/iivision_b200.h declares; the ctypes prototypes cover the header; the
product never imports the oracle."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "iivision_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iiv_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    from iivision_b200 import _build
    return _build.build()


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(built)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n


def test_ctypes_prototypes_cover_header(built):
    from iivision_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared()
    assert _lib.lib.iiv_version() >= 1


def test_host_only_entry_points(built):
    """Calls that touch no device: constants, layout, argument validation."""
    from iivision_b200 import _lib
    assert _lib.mode_info(_lib.MODE_HGR) == (14, 18, 2, [1, 3])
    assert _lib.mode_info(_lib.MODE_DHGR) == (13, 10, 4, [1, 0, 3, 2])
    nbytes, offs = _lib.clip_state_layout()
    assert offs[0] == 0 and offs == sorted(offs) and nbytes > offs[-1]
    assert offs[1] - offs[0] == 32 * 128 * 8 and offs[2] - offs[1] == 8192
    rc = _lib.lib.iiv_pack(7, None, None, 0, None, 0, None)
    assert rc == -1 and b"mode" in _lib.lib.iiv_last_error()
    with pytest.raises(_lib.IIVError):
        _lib.check(rc)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "iivision_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "liboracle" not in text, f


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from iivision_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.table_generate("DHGR", [[0] * 16] * 16)

"""Python face of the table-generator oracle (ctypes over tables_oracle.c).

TEST INFRASTRUCTURE ONLY.  Follows reference transcoder/make_data_tables.py
(:55-174), transcoder/colours.py (:83-148) and transcoder/screen.py (:343-367,
:710-789, :982-990); see tables_oracle.c for the per-function citations and the
"parity unpinned" statement for absolute table values.
"""

import ctypes
import os
import subprocess

import numpy as np

from . import cie2000, palettes

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_tables.so")
_lib = None

MODES = {"HGR": 0, "DHGR": 1}
MASKED_BITS = {"HGR": 14, "DHGR": 13}
MASKED_DOTS = {"HGR": 18, "DHGR": 10}
NUM_OFFSETS = {"HGR": 2, "DHGR": 4}
PHASES = {"HGR": (1, 3), "DHGR": (1, 0, 3, 2)}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tables_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_tables.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_to_dots.restype = ctypes.c_uint32
        L.oracle_to_dots.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_int]
        L.oracle_all_pixel_strings.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.oracle_all_dots.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.oracle_dam_lev.restype = ctypes.c_double
        L.oracle_dam_lev.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int,
                                     dp, dp, dp, dp]
        L.oracle_chain_distance.restype = ctypes.c_int32
        L.oracle_chain_distance.argtypes = [
            u8p, u8p, ctypes.c_int, ctypes.c_void_p]
        L.oracle_build_table.restype = ctypes.c_int64
        L.oracle_build_table.argtypes = [
            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
            ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
        L.oracle_build_table_strided.restype = ctypes.c_int64
        L.oracle_build_table_strided.argtypes = [
            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
            ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
        L.oracle_symmetrise.argtypes = [ctypes.c_int, ctypes.c_void_p]
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def substitution_lut(palette_id: int) -> np.ndarray:
    """int32[16][16] = int(dE2000) (make_data_tables.py:55-70)."""
    return cie2000.diff_matrix(palettes.RGB[palette_id])


def all_dots(mode: str) -> np.ndarray:
    out = np.zeros((NUM_OFFSETS[mode], 1 << MASKED_BITS[mode]), dtype=np.uint32)
    lib().oracle_all_dots(MODES[mode], out.ctypes.data)
    return out


def all_pixel_strings(mode: str) -> np.ndarray:
    out = np.zeros(
        (NUM_OFFSETS[mode], 1 << MASKED_BITS[mode], MASKED_DOTS[mode]),
        dtype=np.uint8)
    lib().oracle_all_pixel_strings(MODES[mode], out.ctypes.data)
    return out


def dam_lev(a: bytes, b: bytes, ins, dele, sub, tr=None) -> float:
    """weighted_levenshtein.dam_lev(a, b, insert_costs, delete_costs,
    substitute_costs[, transpose_costs]) restated (make_data_tables.py:98-104)."""
    u8p = ctypes.POINTER(ctypes.c_uint8)
    dp = ctypes.POINTER(ctypes.c_double)
    ab = (ctypes.c_uint8 * len(a)).from_buffer_copy(a)
    bb = (ctypes.c_uint8 * len(b)).from_buffer_copy(b)
    ins = np.ascontiguousarray(ins, dtype=np.float64)
    dele = np.ascontiguousarray(dele, dtype=np.float64)
    sub = np.ascontiguousarray(sub, dtype=np.float64)
    trp = None
    if tr is not None:
        tr = np.ascontiguousarray(tr, dtype=np.float64)
        trp = tr.ctypes.data_as(dp)
    return lib().oracle_dam_lev(
        ctypes.cast(ab, u8p), len(a), ctypes.cast(bb, u8p), len(b),
        ins.ctypes.data_as(dp), dele.ctypes.data_as(dp),
        sub.ctypes.data_as(dp), trp)


def chain_distance(a: np.ndarray, b: np.ndarray, lut: np.ndarray) -> int:
    u8p = ctypes.POINTER(ctypes.c_uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    lut = np.ascontiguousarray(lut, dtype=np.int32)
    return lib().oracle_chain_distance(
        a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), len(a), lut.ctypes.data)


def build_table(mode: str, lut: np.ndarray, row_begin: int = 0,
                row_end: int = None, faithful: bool = False,
                triangular: bool = True, out: np.ndarray = None,
                threads: int = None, row_step: int = 1):
    """compute_edit_distance restated (make_data_tables.py:111-174).

    Returns (table uint16[n_off, 4**bits], entries evaluated).  ``faithful``
    runs the full (n+2)^2 float64 dam_lev per pair; otherwise the exact 1-D
    recurrence.  Only rows row_begin, row_begin + row_step, ... < row_end are
    filled.
    """
    bits = MASKED_BITS[mode]
    if row_end is None:
        row_end = 1 << bits
    if out is None:
        out = np.zeros((NUM_OFFSETS[mode], 1 << (2 * bits)), dtype=np.uint16)
    lut = np.ascontiguousarray(lut, dtype=np.int32)
    if threads is not None:
        lib().oracle_set_threads(int(threads))
    n = lib().oracle_build_table_strided(
        MODES[mode], lut.ctypes.data, out.ctypes.data, row_begin, row_end,
        max(1, int(row_step)), 0 if faithful else 1, 1 if triangular else 0)
    return out, int(n)


def symmetrise(mode: str, table: np.ndarray) -> np.ndarray:
    """Bitmap.edit_distances' in-memory form (screen.py:358-365), in place."""
    assert table.dtype == np.uint16 and table.flags.c_contiguous
    lib().oracle_symmetrise(MODES[mode], table.ctypes.data)
    return table

"""Opt-in pin of the table generator against the reference's own third-party arithmetic.

The reference computes the 16x16 substitution LUT with colormath 3.0.0 and every table
entry with weighted-levenshtein 0.2.2 (requirements.txt:6,32; call sites
make_data_tables.py:63-69 and :98-104).  Neither wheel is available offline, so these tests
skip unless both import; anyone with the reference's environment pins the generator with

    pip install colormath==3.0.0 weighted-levenshtein==0.2.2
    python -m pytest tests/test_reference_env.py -m reference_env -q -s

The oracle half needs no GPU; the CUDA half reports the count of table entries that differ
from the reference by exactly one unit (north star: fewer than 1e-4 of entries) and fails
on any larger difference.
"""

import numpy as np
import pytest

pytestmark = pytest.mark.reference_env


@pytest.fixture(scope="module")
def reference_libs():
    color_conversions = pytest.importorskip("colormath.color_conversions")
    color_objects = pytest.importorskip("colormath.color_objects")
    color_diff = pytest.importorskip("colormath.color_diff")
    wl = pytest.importorskip("weighted_levenshtein")
    return color_conversions, color_objects, color_diff, wl


def _reference_diff_matrix(libs, rgb16, as_float=False):
    """compute_diff_matrix (make_data_tables.py:55-70) on a uint8[16][3] palette."""
    conv, objs, diff, _ = libs
    labs = [conv.convert_color(objs.sRGBColor(*(int(c) for c in rgb), is_upscaled=True),
                               objs.LabColor) for rgb in rgb16]
    dm = np.zeros((16, 16), dtype=np.float64)
    for i, a in enumerate(labs):
        for j, b in enumerate(labs):
            dm[i, j] = diff.delta_e_cie2000(a, b)
    return dm if as_float else dm.astype(np.int32)    # int() truncates, values >= 0


@pytest.mark.parametrize("pid", [0, 5])
def test_oracle_lut_equals_colormath(reference_libs, pid):
    from oracle import cie2000, palettes
    want = _reference_diff_matrix(reference_libs, palettes.RGB[pid])
    got = cie2000.diff_matrix(palettes.RGB[pid])
    wf = _reference_diff_matrix(reference_libs, palettes.RGB[pid], as_float=True)
    gf = cie2000.diff_matrix_float(palettes.RGB[pid])
    print("palette %d: max |dE oracle - dE colormath| = %.3g, integer mismatches %d / 256"
          % (pid, np.abs(wf - gf).max(), int((want != got).sum())))
    assert np.array_equal(got, want)


def _random_pairs(n_pairs, dots, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 16, size=(n_pairs, dots), dtype=np.uint8)
    b = a.copy()
    # realistic pairs: mostly a few substitutions, some adjacent swaps
    flips = rng.random((n_pairs, dots)) < 0.3
    b[flips] = rng.integers(0, 16, size=int(flips.sum()), dtype=np.uint8)
    for k in range(0, n_pairs, 3):
        t = int(rng.integers(0, dots - 1))
        b[k, t], b[k, t + 1] = a[k, t + 1], a[k, t]
    return a, b


def _reference_dam_lev(libs, lut, a, b):
    """edit_distance (make_data_tables.py:92-108) with EditDistanceParams' costs."""
    wl = libs[3]
    chars = "0123456789ABCDEF"
    ins = np.ones(128, dtype=np.float64) * 100000
    dele = np.ones(128, dtype=np.float64) * 100000
    sub = np.zeros((128, 128), dtype=np.float64)
    for i, c in enumerate(chars):
        for j, d in enumerate(chars):
            sub[ord(c), ord(d)] = lut[i, j]
            sub[ord(d), ord(c)] = lut[i, j]
    out = np.zeros(len(a), dtype=np.float64)
    for k in range(len(a)):
        sa = "".join(chars[v] for v in a[k])
        sb = "".join(chars[v] for v in b[k])
        out[k] = wl.dam_lev(sa, sb, insert_costs=ins, delete_costs=dele, substitute_costs=sub)
    return out


@pytest.mark.parametrize("dots", [18, 10])
def test_oracle_dam_lev_equals_weighted_levenshtein(reference_libs, dots):
    from oracle import palettes, tables
    lut = _reference_diff_matrix(reference_libs, palettes.RGB[5])
    a, b = _random_pairs(20000, dots, seed=dots)
    want = _reference_dam_lev(reference_libs, lut, a, b)
    got = np.array([tables.chain_distance(a[k], b[k], lut) for k in range(len(a))])
    assert np.array_equal(got, want.astype(np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("pid", [0, 5])
def test_cuda_lut_and_dam_lev_against_reference(reference_libs, pid):
    from iivision_b200 import ops
    from oracle import palettes
    want_lut = _reference_diff_matrix(reference_libs, palettes.RGB[pid])
    got_lut = ops.lut_cie2000(palettes.RGB[pid])
    assert np.array_equal(np.asarray(got_lut), want_lut)
    total = off_by_one = 0
    for dots in (18, 10):
        a, b = _random_pairs(50000, dots, seed=100 + dots)
        want = _reference_dam_lev(reference_libs, want_lut, a, b)
        got = np.asarray(ops.string_distance(got_lut, a, b)).astype(np.int64)
        d = np.abs(got - want.astype(np.int64))
        assert d.max() <= 1
        total += len(d)
        off_by_one += int((d == 1).sum())
    print("palette %d: %d of %d dam_lev results differ from weighted-levenshtein by one unit"
          % (pid, off_by_one, total))
    assert off_by_one < 1e-4 * total

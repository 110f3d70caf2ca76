"""CPU: the table-generator oracle against every known answer we hold.

Absolute table values are "parity unpinned" (colormath / weighted-levenshtein are
not vendored); what IS pinned here: pixel strings against hashes
produced by the reference code (tests/golden/pixel_strings.json), the CIEDE2000
restatement against Sharma et al.'s published test data, the LUTs against the
rows recorded in SURVEY.md Appendix C, and the 1-D chain against the restated
full Damerau-Levenshtein DP.
"""

import hashlib
import json
import os

import numpy as np
import pytest

from oracle import cie2000, palettes, tables

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_ciede2000_sharma_table1():
    for p in _json("ciede2000_sharma.json")["pairs"]:
        assert abs(cie2000.delta_e_cie2000(p[:3], p[3:6]) - p[6]) < 1e-4, p


@pytest.mark.parametrize("pid", [0, 5])
def test_lut_known_answers(oracle_luts, pid):
    g = _json("luts.json")
    lut = oracle_luts[pid]
    assert lut.tolist() == g["lut"][str(pid)]
    assert lut[0].tolist() == g["survey_row0"][str(pid)]
    assert np.array_equal(lut, lut.T) and not np.diagonal(lut).any()
    # black <-> white sits 1.5e-5 under 100: int() gives 99
    f = cie2000.diff_matrix_float(palettes.RGB[pid])
    assert lut[0, 15] == 99 and 99.9999 < f[0, 15] < 100.0
    if pid == 5:
        assert lut[5, 10] == 0 and lut.max() == 101
        assert lut[lut > 0].min() == 14 and lut[12, 14] == 14
    else:
        assert lut.max() == 110 and lut[lut > 0].min() == 13
    # every other entry is far from an integer boundary
    frac = np.minimum(f - np.floor(f), np.ceil(f) - f)
    frac[0, 15] = frac[15, 0] = 1
    frac[np.arange(16), np.arange(16)] = 1
    if pid == 5:
        frac[5, 10] = frac[10, 5] = 1
    assert frac.min() > 1e-3


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_pixel_strings_match_reference_hashes(mode):
    g = _json("pixel_strings.json")[mode]
    dots = tables.all_dots(mode)
    pix = tables.all_pixel_strings(mode)
    assert list(pix.shape) == g["shape"]
    assert _sha(dots) == g["dots_sha256"]
    assert _sha(pix) == g["pixels_sha256"]
    assert int(pix.sum()) == g["pixels_sum"]


def test_survey_hashes():
    """The hashes recorded independently in SURVEY.md 8(c)."""
    assert _sha(tables.all_pixel_strings("HGR")) == \
        "9761a47a25c8c6e3e902fc7e942bf8068ac20d64ceaf4498ef0500a2a544fc22"
    assert _sha(tables.all_pixel_strings("DHGR")) == \
        "9f8823a8e60c95a80687c078047b6a77a62ab29676c3ec45ca1ad690e0f660c1"
    assert _sha(tables.all_dots("HGR")) == \
        "8f102cf1d95a8b1406497a10c12c813d92ebe9f61dfb4ff390e5b20db52731ab"
    assert _sha(tables.all_dots("DHGR")) == \
        "86a793c62ab6d02292f2b16f073ad02716dd1f0788117c328c0ef1bb658e6fee"


def _unit_costs():
    return np.ones(128), np.ones(128), np.ones((128, 128)), np.ones((128, 128))


def test_dam_lev_classic_answers():
    ins, dele, sub, tr = _unit_costs()
    for a, b, want in ((b"", b"", 0), (b"abc", b"abc", 0), (b"ab", b"ba", 1),
                       (b"kitten", b"sitting", 3), (b"ca", b"abc", 2),
                       (b"abcdef", b"", 6), (b"a", b"b", 1), (b"abcd", b"acbd", 1)):
        assert tables.dam_lev(a, b, ins, dele, sub, tr) == want, (a, b)
        assert tables.dam_lev(a, b, ins, dele, sub) == want      # default transpose = 1


def test_chain_equals_full_dp_random():
    """with insert/delete at 1e5 the (n+2)^2 DP collapses to the chain."""
    rng = np.random.default_rng(0)
    digits = b"0123456789ABCDEF"
    for trial in range(3000):
        n = (10, 18)[trial & 1]
        k = (2, 3, 16)[trial % 3]
        lut = rng.integers(0, 120, size=(16, 16)).astype(np.int32)
        lut = np.maximum(lut, lut.T)
        lut[rng.random((16, 16)) < 0.1] = 0
        lut = np.minimum(lut, lut.T)
        np.fill_diagonal(lut, 0)
        a = rng.integers(0, k, size=n).astype(np.uint8)
        b = rng.integers(0, k, size=n).astype(np.uint8)
        ins = np.full(128, 1e5)
        sub = np.zeros((128, 128))
        for i in range(16):
            for j in range(16):
                sub[digits[i], digits[j]] = lut[i, j]
        full = tables.dam_lev(bytes(digits[x] for x in a), bytes(digits[x] for x in b),
                              ins, ins, sub)
        assert full == tables.chain_distance(a, b, lut)


@pytest.mark.parametrize("mode", ["DHGR", "HGR"])
def test_chain_equals_faithful_on_table_rows(oracle_luts, mode):
    n = 1 << tables.MASKED_BITS[mode]
    for begin in (1, n // 3, n - 4):
        fast, c1 = tables.build_table(mode, oracle_luts[5], begin, begin + 4)
        slow, c2 = tables.build_table(mode, oracle_luts[5], begin, begin + 4,
                                      faithful=True)
        assert c1 == c2 and np.array_equal(fast, slow)


@pytest.mark.parametrize("mode,pid", [("DHGR", 5), ("DHGR", 0), ("HGR", 5), ("HGR", 0)])
def test_table_hash_and_layout(oracle_luts, mode, pid):
    tab, n = tables.build_table(mode, oracle_luts[pid])
    assert _sha(tab) == _json("luts.json")["table_sha256"]["%s_%d" % (mode, pid)]
    bits = tables.MASKED_BITS[mode]
    assert n == tables.NUM_OFFSETS[mode] * (1 << bits) * ((1 << bits) - 1) // 2
    # lower-triangular file layout (make_data_tables.py:156-172): rows i hold j < i
    sq = tab[0].reshape(1 << bits, 1 << bits)
    assert not np.triu(sq[:512, :512]).any()


def test_dhgr_ntsc_known_answers(oracle_tables):
    """max 1010, T[0][0,0x1FFF] = 990, two off-diagonal zeros per
    offset (GREY1 == GREY2 in the NTSC palette)."""
    t = oracle_tables("DHGR", 5)
    assert t.max() == 1010 and t[0][0x1FFF] == 990
    for o in range(4):
        sq = t[o].reshape(8192, 8192)
        assert np.array_equal(sq, sq.T)
        zi, zj = np.nonzero(sq == 0)
        off = sorted((int(i), int(j)) for i, j in zip(zi, zj) if i != j)
        assert off == [(0x0AAA, 0x1555), (0x1555, 0x0AAA)]


def test_hgr_ntsc_known_answers(oracle_tables):
    t = oracle_tables("HGR", 5)
    assert t.max() == 1818
    sq = t[1].reshape(16384, 16384)
    assert np.array_equal(sq[:2048, :2048], sq[:2048, :2048].T)


@pytest.mark.parametrize("mode,leaf", [("HGR", 4), ("DHGR", 3)])
def test_low_bits_reach_leaf_pixels_only(mode, leaf):
    """The structural fact the tree generator (csrc/iiv_tables.cu) relies on: the low
    three bits of a masked value only influence the first `leaf` pixels, and in the
    pattern the kernel's tree walks."""
    pix = tables.all_pixel_strings(mode)
    v = np.arange(1 << tables.MASKED_BITS[mode])
    n = tables.MASKED_DOTS[mode]
    want = ({0: {0, 1}, 1: {0, 1, 2, 3}, 2: {0, 1}} if mode == "HGR"
            else {0: {0}, 1: {0, 1}, 2: {0, 1, 2}})
    for o in range(pix.shape[0]):
        for bit in range(3):
            dep = {t for t in range(n) if (pix[o][v, t] != pix[o][v ^ (1 << bit), t]).any()}
            assert dep <= want[bit] and max(dep) < leaf, (mode, o, bit, dep)


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_split_windows(mode):
    """The structural fact the split generator (csrc/iiv_tables.cu, ALGO_SPLIT) relies on:
    pixels 0..cut of a value's string are functions of the bits in mask_a only, pixels
    cut..n-1 of the bits in mask_b only -- the windows the library reports (host-only
    call) against the oracle's pixel strings."""
    import ctypes
    from iivision_b200 import _lib
    pix = tables.all_pixel_strings(mode)
    bits, n = tables.MASKED_BITS[mode], tables.MASKED_DOTS[mode]
    v = np.arange(1 << bits)
    for o in range(pix.shape[0]):
        cut, ma, mb = ctypes.c_int(), ctypes.c_uint32(), ctypes.c_uint32()
        _lib.check(_lib.lib.iiv_table_split_windows(
            tables.MODES[mode], o,
            ctypes.byref(cut), ctypes.byref(ma), ctypes.byref(mb)))
        cut, ma, mb = cut.value, ma.value, mb.value
        assert 0 < cut < n - 1 and ma < (1 << bits) and mb < (1 << bits)
        assert np.array_equal(pix[o][v, :cut + 1], pix[o][v & ma, :cut + 1]), (mode, o)
        assert np.array_equal(pix[o][v, cut:], pix[o][v & mb, cut:]), (mode, o)
        # what the kernel's thread mapping assumes about the windows
        assert ma & 0xff == 0xff and mb & 7 == 0
        assert bin(ma).count("1") + 5 == bits
    with pytest.raises(_lib.IIVError):
        _lib.check(_lib.lib.iiv_table_split_windows(0, 2, ctypes.byref(ctypes.c_int()),
                                                    ctypes.byref(ctypes.c_uint32()),
                                                    ctypes.byref(ctypes.c_uint32())))


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_factor_segments(mode):
    """The structural facts the factored scorer (csrc/iiv_factored.cu) relies on: its pixel
    segments tile 0..n, and pixels p..min(q, n-1) of a value's string are functions of the
    segment's bit window only -- against the oracle's pixel strings.  Also that the
    segment-by-segment (min,+) product equals the oracle's chain on random pairs."""
    import ctypes
    from iivision_b200 import _lib
    pix = tables.all_pixel_strings(mode).astype(np.int64)
    bits, n = tables.MASKED_BITS[mode], tables.MASKED_DOTS[mode]
    v = np.arange(1 << bits)
    rng = np.random.default_rng(5)
    lut = tables.substitution_lut(0).astype(np.int64)
    for o in range(pix.shape[0]):
        cnt = ctypes.c_int()
        p, q, m = (ctypes.c_int * 16)(), (ctypes.c_int * 16)(), (ctypes.c_uint32 * 16)()
        _lib.check(_lib.lib.iiv_score_factor_segments(tables.MODES[mode], o, ctypes.byref(cnt),
                                                      p, q, m))
        segs = [(p[k], q[k], m[k]) for k in range(cnt.value)]
        assert segs[0][0] == 0 and segs[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(segs[:-1], segs[1:]))
        room = 0
        for k, (sp, sq, mask) in enumerate(segs):
            hi = min(sq, n - 1)
            assert np.array_equal(pix[o][v, sp:hi + 1], pix[o][v & mask, sp:hi + 1]), (mode, o, k)
            room += (4 if k in (0, len(segs) - 1) else 8) << (2 * bin(mask).count("1"))
        assert 2 * room <= 227 * 1024          # a bank's two offsets in one SM's shared memory
        assert room * pix.shape[0] + 16 == _lib.lib.iiv_score_factors_bytes(tables.MODES[mode])
        i, j = rng.integers(0, 1 << bits, 2000), rng.integers(0, 1 << bits, 2000)
        want = np.array([tables.chain_distance(pix[o][a], pix[o][b], lut) for a, b in zip(i, j)])
        f1, f2 = np.zeros(2000, np.int64), np.full(2000, 0x4000, np.int64)
        for sp, sq, mask in reversed(segs):
            a, b = pix[o][i & mask], pix[o][j & mask]
            for t in range(sq - 1, sp - 1, -1):
                cur = f1 + lut[a[:, t], b[:, t]]
                if t + 1 < n:
                    sw = (a[:, t] == b[:, t + 1]) & (a[:, t + 1] == b[:, t])
                    cur = np.where(sw, np.minimum(cur, f2 + 1), cur)
                f2, f1 = f1, cur
        assert np.array_equal(f1, want), (mode, o)

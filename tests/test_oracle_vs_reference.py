"""CPU, build container only: the oracle against the UNMODIFIED reference modules
imported from /root/reference (skipped where the tree is not mounted, e.g. on the
GPU box -- tests/golden carries the same evidence there)."""

import random

import numpy as np
import pytest

from oracle import ref_harness, scorer, tables

pytestmark = [
    pytest.mark.reference,
    pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted"),
    pytest.mark.filterwarnings("ignore::RuntimeWarning"),
]


@pytest.fixture(scope="module")
def ns():
    return ref_harness.load()


def _ref_bitmap(ns, mode, main, aux=None):
    mm = ns.screen.MemoryMap(screen_page=1, page_offset=main)
    if mode == "DHGR":
        am = ns.screen.MemoryMap(screen_page=1, page_offset=aux)
        return ns.screen.DHGRBitmap(palette=ns.palette.Palette.NTSC, main_memory=mm,
                                    aux_memory=am)
    return ns.screen.HGRBitmap(palette=ns.palette.Palette.NTSC, main_memory=mm)


def test_constants_and_holes(ns):
    assert np.array_equal(scorer.SCREEN_HOLES, ns.screen.SCREEN_HOLES)
    for mode, cls in (("HGR", ns.screen.HGRBitmap), ("DHGR", ns.screen.DHGRBitmap)):
        spec = scorer.SPECS[mode]
        assert spec.masked_bits == int(cls.MASKED_BITS)
        assert tables.MASKED_DOTS[mode] == int(cls.MASKED_DOTS)
        assert list(tables.PHASES[mode]) == list(cls.PHASES)
        assert [int(m) for m in spec.masks] == [int(m) for m in cls.BYTE_MASKS]
        assert [int(s) for s in spec.shifts] == [int(s) for s in cls.BYTE_SHIFTS]
        for off in range(256):
            for is_aux in ((False, True) if mode == "DHGR" else (False,)):
                assert spec.byte_offset(off, is_aux) == cls.byte_offset(off, is_aux)


def test_palettes(ns):
    from oracle import palettes
    for pid, cls in ((5, ns.palette.NTSCPalette), (0, ns.palette.IIGSPalette)):
        assert cls.ID.value == pid
        for colour, rgb in cls.RGB.items():
            got = (rgb.rgb_r, rgb.rgb_g, rgb.rgb_b)
            assert tuple(int(x) for x in palettes.RGB[pid][colour.value]) == tuple(
                int(round(c * 255)) if isinstance(c, float) and c <= 1.0 and rgb.is_upscaled is False
                else int(c) for c in got)


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_to_dots_and_pixels_sample(ns, mode):
    cls = ns.screen.HGRBitmap if mode == "HGR" else ns.screen.DHGRBitmap
    cols = ns.colours.HGRColours if mode == "HGR" else ns.colours.DHGRColours
    dots = tables.all_dots(mode)
    pix = tables.all_pixel_strings(mode)
    rng = np.random.default_rng(1)
    n = tables.MASKED_DOTS[mode]
    for o, ph in enumerate(cls.PHASES):
        for v in rng.integers(0, 1 << tables.MASKED_BITS[mode], size=600):
            d = cls.to_dots(int(v), o)
            assert d == dots[o, v]
            want = ns.colours.dots_to_nominal_colour_pixel_values(n, d, cols, init_phase=ph)
            assert tuple(pix[o, v]) == tuple(want)


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_random_screens(ns, oracle_tables, mode):
    table = oracle_tables(mode)
    ref_harness.install_tables(ns, mode, {5: table})
    rng = np.random.default_rng(3)
    hi = 256 if mode == "HGR" else 128
    for trial in range(3):
        mems = [rng.integers(0, hi, size=(32, 256), dtype=np.uint8) for _ in range(4)]
        for m in mems:
            m[scorer.SCREEN_HOLES] = 0
        rs = _ref_bitmap(ns, mode, mems[0].copy(), mems[1].copy())
        rt = _ref_bitmap(ns, mode, mems[2].copy(), mems[3].copy())
        aux = (lambda k: mems[k].copy()) if mode == "DHGR" else (lambda k: None)
        os_ = scorer.OracleBitmap(mode, table, mems[0].copy(), aux(1))
        ot = scorer.OracleBitmap(mode, table, mems[2].copy(), aux(3))
        assert np.array_equal(os_.packed, rs.packed)
        for is_aux in ((False, True) if mode == "DHGR" else (False,)):
            dw_r = rt.diff_weights(rs, is_aux)
            dw_o = ot.diff_weights(os_, is_aux)
            assert np.array_equal(dw_r, dw_o)
            for _ in range(20):
                page, content = int(rng.integers(0, 32)), np.uint8(rng.integers(0, hi))
                assert np.array_equal(
                    rt.compute_delta_page(page, content, dw_r[page, :], is_aux),
                    ot.compute_delta_page(page, content, dw_o[page, :], is_aux))
        for _ in range(500):
            page, off = int(rng.integers(0, 32)), int(rng.integers(0, 256))
            is_aux = bool(rng.integers(0, 2)) if mode == "DHGR" else False
            val = np.uint8(rng.integers(0, hi))
            rs.apply(page, off, is_aux, val)
            os_.apply(page, off, is_aux, val)
        assert np.array_equal(os_.packed, rs.packed)
        assert np.array_equal(os_.main, rs.main_memory.page_offset)


@pytest.mark.parametrize("mode,fraction", [("DHGR", 0.3), ("HGR", 0.3)])
def test_live_encode_run(ns, oracle_tables, mode, fraction):
    """A fresh encode run (not one of the committed fixtures) through both."""
    from iivision_b200 import synth
    table = oracle_tables(mode)
    ref_harness.install_tables(ns, mode, {5: table})
    frames = synth.synthetic_frames(mode, 2, fraction, seed=77)
    segs = synth.movie_schedule(mode, 2, 500, 170)
    vm = getattr(ns.video_mode.VideoMode, mode)
    random.seed(5)
    np.random.seed(5)
    rv = ns.video.Video(ns.frame_grabber.FrameGrabber(vm), ticks_per_second=14700.,
                        mode=vm, palette=ns.palette.Palette.NTSC)
    ov = scorer.OracleVideo(mode, table, py_rng=random.Random(5),
                            np_rng=np.random.RandomState(5))
    for frame, is_aux, budget in segs:
        aux = frames[frame, 1] if mode == "DHGR" else None
        rt = _ref_bitmap(ns, mode, frames[frame, 0].copy(), None if aux is None else aux.copy())
        ot = ov.target_bitmap(frames[frame, 0], aux)
        rseq, oseq = rv.encode_frame(rt, bool(is_aux)), ov.encode_frame(ot, bool(is_aux))
        for _ in range(budget):
            rp, rc, ro = next(rseq)
            op, oc, oo = next(oseq)
            assert (int(rp), int(rc), [int(x) for x in ro]) == (int(op), int(oc), [int(x) for x in oo])
    assert np.array_equal(rv.pixelmap.packed, ov.pixelmap.packed)
    assert np.array_equal(rv.update_priority, ov.update_priority)

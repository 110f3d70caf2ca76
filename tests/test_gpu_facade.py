"""GPU: the reference-named Python facades (make_data_tables / screen / video)
against fixtures produced by the unmodified reference and against the oracle."""

import contextlib
import io
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def mods():
    from iivision_b200 import colours, make_data_tables, palette, screen, video, video_mode
    import types
    return types.SimpleNamespace(colours=colours, mdt=make_data_tables, palette=palette,
                                 screen=screen, video=video, video_mode=video_mode)


class _Grabber:
    input_frame_rate = 30


def test_make_data_tables_facade(mods, oracle_luts):
    from oracle import tables
    for pal in (mods.palette.NTSCPalette, mods.palette.IIGSPalette):
        pid = pal.ID.value
        assert np.array_equal(mods.mdt.compute_diff_matrix(pal), oracle_luts[pid])
        edp = mods.mdt.compute_substitute_costs(pal)
        assert edp.substitute_costs[ord("0"), ord("F")] == 99
        assert edp.error_substitute_costs[ord("3"), ord("C")] == 5 * oracle_luts[pid][3, 12]
        assert edp.substitute_costs[ord("A"), ord("a")] == 0       # only '0'..'F' filled
    edp = mods.mdt.compute_substitute_costs(mods.palette.NTSCPalette)
    got = mods.mdt.compute_edit_distance(edp, mods.screen.DHGRBitmap, mods.colours.DHGRColours)
    want, _ = tables.build_table("DHGR", oracle_luts[5], triangular=True)
    assert got.dtype == np.uint16 and got.shape == want.shape
    assert np.array_equal(got, want)
    # edit_distance on explicit strings, both cost sets (make_data_tables.py:92-108)
    a, b = "0123456789", "1023456798"
    lut = oracle_luts[5]
    na = np.array([int(c, 16) for c in a], np.uint8)
    nb = np.array([int(c, 16) for c in b], np.uint8)
    assert mods.mdt.edit_distance(edp, a, b, error=False) == tables.chain_distance(na, nb, lut)
    assert mods.mdt.edit_distance(edp, a, b, error=True) == tables.chain_distance(na, nb, 5 * lut)
    assert mods.mdt.pixel_string((1, 2, 15)) == "12F"
    # parameter sets the collapse does not cover are refused, not mis-computed
    bad = mods.mdt.EditDistanceParams()
    saved = bad.insert_costs.copy()
    try:
        type(bad).insert_costs[:] = 10
        with pytest.raises(ValueError):
            mods.mdt.compute_edit_distance(bad, mods.screen.DHGRBitmap)
    finally:
        type(bad).insert_costs[:] = saved


def test_make_edit_distance_writes_reference_file(mods, oracle_luts, tmp_path, monkeypatch):
    """npz name/key/layout as the reference writes them; the loader symmetrises it."""
    from oracle import tables
    monkeypatch.chdir(tmp_path)
    os.makedirs("transcoder/data")
    pal = mods.palette.IIGSPalette
    edp = mods.mdt.compute_substitute_costs(pal)
    mods.mdt.make_edit_distance(pal, edp, mods.screen.DHGRBitmap, mods.colours.DHGRColours)
    path = "transcoder/data/DHGR_palette_0_edit_distance.npz"
    tri = np.load(path)["edit_distance"]
    want, _ = tables.build_table("DHGR", oracle_luts[0], triangular=True)
    assert np.array_equal(tri, want)
    mods.screen.DHGRBitmap.edit_distances_device.cache_clear()
    mods.screen.DHGRBitmap.edit_distances.cache_clear()
    try:
        sym = mods.screen.DHGRBitmap.edit_distances(mods.palette.Palette.IIGS)   # from the file
        want_sym = tables.symmetrise("DHGR", want)      # in place
        assert np.array_equal(sym, want_sym)
        # and from a file as the reference's own generator writes it (numpy's writer)
        np.savez_compressed(path, edit_distance=tri)
        mods.screen.DHGRBitmap.edit_distances_device.cache_clear()
        mods.screen.DHGRBitmap.edit_distances.cache_clear()
        sym = mods.screen.DHGRBitmap.edit_distances(mods.palette.Palette.IIGS)
        assert np.array_equal(sym, want_sym)
    finally:
        mods.screen.DHGRBitmap.edit_distances_device.cache_clear()
        mods.screen.DHGRBitmap.edit_distances.cache_clear()


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_screen_facade(mods, mode):
    g = np.load(os.path.join(GOLDEN, "scorer_%s.npz" % mode.lower()))
    s = mods.screen
    fr = g["frames"]

    def bitmap(k):
        mm = s.MemoryMap(screen_page=1, page_offset=fr[k, 0].copy())
        if mode == "DHGR":
            return s.DHGRBitmap(palette=mods.palette.Palette.NTSC, main_memory=mm,
                                aux_memory=s.MemoryMap(1, fr[k, 1].copy()))
        return s.HGRBitmap(palette=mods.palette.Palette.NTSC, main_memory=mm)
    src, tgt = bitmap(0), bitmap(1)
    cls = type(src)
    assert np.array_equal(src.packed, g["src_packed"])
    assert np.array_equal(tgt.packed, g["tgt_packed"])
    for o in range(len(cls.BYTE_MASKS)):
        assert np.array_equal(cls.mask_and_shift_data(g["words"], o), g["mask_shift"][o])
        assert cls.mask_and_shift_data(g["words"][0, 0], o) == g["mask_shift"][o][0, 0]
        for k, v in enumerate(g["values"]):
            assert np.array_equal(cls.masked_update(o, g["words"], np.uint8(v)),
                                  g["masked_update"][o, k])
    for tag in (("main", "aux") if mode == "DHGR" else ("main",)):
        is_aux = tag == "aux"
        dw = tgt.diff_weights(src, is_aux)
        assert dw.dtype == np.int32 and np.array_equal(dw, g["diff_weights_" + tag])
        for (page, content), want in zip(g["delta_cases_" + tag], g["delta_" + tag]):
            got = tgt.compute_delta_page(int(page), np.uint8(content), dw[page, :], is_aux)
            assert np.array_equal(got, want)
        for bo, page, off, content, want in g["pair_difference_" + tag]:
            assert int(tgt.byte_pair_difference(int(bo), tgt.packed[page, off // 2],
                                                np.uint8(content))) == want
    for page, off, is_aux, val in g["apply_stores"][:40]:
        src.apply(int(page), int(off), bool(is_aux), np.uint8(val))
    src.apply_many([tuple(int(x) for x in st) for st in g["apply_stores"][40:]])
    assert np.array_equal(src.packed, g["apply_packed"])
    assert np.array_equal(src.main_memory.page_offset, g["apply_main"])
    src._check_consistency()
    assert np.array_equal(s.SCREEN_HOLES.sum(), 512)
    with pytest.raises(ValueError):
        s.MemoryMap(screen_page=3)
    with pytest.raises(ValueError):
        s.MemoryMap(screen_page=1, page_offset=np.zeros((3, 3), np.uint8))


@pytest.mark.parametrize("name", ["dhgr_sparse", "hgr_full"])
@pytest.mark.parametrize("speculate", [None, 100])
def test_video_facade_streams(mods, name, speculate):
    """Video.encode_frame pulled like Movie.encode pulls it (new generator per segment,
    old one abandoned), against the reference's own opcode stream."""
    g = np.load(os.path.join(GOLDEN, "stream_%s.npz" % name))
    mode = str(g["mode"])
    seed = int(g["rng_seed"])
    vm = getattr(mods.video_mode.VideoMode, mode)
    random.seed(seed)
    np.random.seed(seed)
    v = mods.video.Video(_Grabber(), ticks_per_second=14700., mode=vm,
                         palette=mods.palette.Palette.NTSC, speculate=speculate)
    frames = g["frames"]
    got, sims = [], []
    op_seq = None
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        for frame, is_aux, budget in g["segments"]:
            mm = mods.screen.MemoryMap(1, frames[frame, 0].copy())
            if mode == "DHGR":
                tgt = mods.screen.DHGRBitmap(palette=mods.palette.Palette.NTSC, main_memory=mm,
                                             aux_memory=mods.screen.MemoryMap(1, frames[frame, 1].copy()))
            else:
                tgt = mods.screen.HGRBitmap(palette=mods.palette.Palette.NTSC, main_memory=mm)
            op_seq = v.encode_frame(tgt, is_aux=bool(is_aux))     # drops the previous generator
            for _ in range(budget):
                page, content, offs = next(op_seq)
                got.append([page, content] + list(offs))
        op_seq.close()
    assert np.array_equal(np.array(got, np.uint8), g["opcodes"])
    sims = [float(line.split()[1]) for line in out.getvalue().splitlines()]
    assert np.allclose(sims, g["similarity"], atol=1e-6)
    assert np.array_equal(v.pixelmap.packed, g["packed"])
    assert np.array_equal(v.memory_map.page_offset, g["main"])
    assert np.array_equal(v.update_priority, g["priority_main"])
    if mode == "DHGR":
        assert np.array_equal(v.aux_memory_map.page_offset, g["aux"])
        assert np.array_equal(v.aux_update_priority, g["priority_aux"])
    # the process-global generators sit exactly where the reference left them
    assert [random.getrandbits(32) for _ in range(4)] == g["next_python_words"].tolist()
    assert np.random.randint(0, 256, size=4).tolist() == g["next_numpy_bytes"].tolist()


def test_video_sync_mid_generator(mods, oracle_tables):
    """sync() materialises the state after exactly the opcodes pulled so far."""
    from iivision_b200 import synth
    from oracle import scorer
    frames = synth.synthetic_frames("DHGR", 1, 0.4, seed=9)
    random.seed(3)
    np.random.seed(3)
    v = mods.video.Video(_Grabber(), 14700., mode=mods.video_mode.VideoMode.DHGR)
    ov = scorer.OracleVideo("DHGR", oracle_tables("DHGR"), py_rng=random.Random(3),
                            np_rng=np.random.RandomState(3))
    tgt = mods.screen.DHGRBitmap(palette=mods.palette.Palette.NTSC,
                                 main_memory=mods.screen.MemoryMap(1, frames[0, 0].copy()),
                                 aux_memory=mods.screen.MemoryMap(1, frames[0, 1].copy()))
    otgt = ov.target_bitmap(frames[0, 0], frames[0, 1])
    with contextlib.redirect_stdout(io.StringIO()):
        seq, oseq = v.encode_frame(tgt, False), ov.encode_frame(otgt, False)
        for k in range(37):
            a, b = next(seq), next(oseq)
            assert (a[0], a[1], a[2]) == (int(b[0]), int(b[1]), [int(x) for x in b[2]])
        v.sync()
        assert np.array_equal(v.memory_map.page_offset, ov.main)
        assert np.array_equal(v.pixelmap.packed, ov.pixelmap.packed)
        assert np.array_equal(v.update_priority, ov.update_priority)
        for k in range(400):       # beyond the speculated budget: re-run with doubling
            a, b = next(seq), next(oseq)
            assert (a[0], a[1], a[2]) == (int(b[0]), int(b[1]), [int(x) for x in b[2]])
        seq.close()
    assert np.array_equal(v.memory_map.page_offset, ov.main)
    assert np.array_equal(v.update_priority, ov.update_priority)
    assert v.tick(0) and not v.tick(1) and v.tick(490)


def test_video_budget_prediction(mods, oracle_tables, monkeypatch):
    """Pulled the way Movie.encode pulls (one target per frame, a new generator per bank
    flip and per frame), the facade learns both periods: after the first frame every
    generator is one kernel run of exactly the opcodes pulled -- and the stream still
    equals the oracle's."""
    from iivision_b200 import synth
    from oracle import scorer
    n_frames = 3
    frames = synth.synthetic_frames("DHGR", n_frames, 0.5, seed=21)
    segs = synth.movie_schedule("DHGR", n_frames)
    random.seed(5)
    np.random.seed(5)
    v = mods.video.Video(_Grabber(), 14700., mode=mods.video_mode.VideoMode.DHGR)
    ov = scorer.OracleVideo("DHGR", oracle_tables("DHGR"), py_rng=random.Random(5),
                            np_rng=np.random.RandomState(5))
    runs = []
    real_launch = mods.video.Video._launch
    monkeypatch.setattr(
        mods.video.Video, "_launch",
        lambda self, base, tgt, is_aux, budget: (
            runs.append(budget), real_launch(self, base, tgt, is_aux, budget))[1])
    tgt = otgt = None
    tgt_frame = -1
    first_frame_runs = None
    with contextlib.redirect_stdout(io.StringIO()):
        for frame, is_aux, budget in segs:
            if frame != tgt_frame:
                if frame == 1:
                    first_frame_runs = len(runs)
                tgt_frame = frame
                tgt = mods.screen.DHGRBitmap(
                    palette=mods.palette.Palette.NTSC,
                    main_memory=mods.screen.MemoryMap(1, frames[frame, 0].copy()),
                    aux_memory=mods.screen.MemoryMap(1, frames[frame, 1].copy()))
                otgt = ov.target_bitmap(frames[frame, 0], frames[frame, 1])
            seq, oseq = v.encode_frame(tgt, bool(is_aux)), ov.encode_frame(otgt, bool(is_aux))
            for _ in range(budget):
                a, b = next(seq), next(oseq)
                assert (a[0], a[1], a[2]) == (int(b[0]), int(b[1]), [int(x) for x in b[2]])
            seq.close()
    # (with the pipeline on, a generator's launch is issued while its predecessor is still
    # being pulled; the order and the budgets of the launches are the schedule's all the same)
    later = [s_ for s_ in segs if s_[0] >= 1]
    assert runs[first_frame_runs:] == [s_[2] for s_ in later]
    assert np.array_equal(v.update_priority, ov.update_priority)
    assert np.array_equal(v.aux_update_priority, ov.aux_update_priority)
    assert np.array_equal(v.pixelmap.packed, ov.pixelmap.packed)


@pytest.mark.parametrize("name", ["hgr_long_generator", "hgr_exhaust"])
def test_video_facade_long_generators(mods, name):
    """One generator pulled 2450 / 7000 times (HGR never flips banks; the reference's loop
    is unbounded, video.py:121), against the reference's own stream."""
    g = np.load(os.path.join(GOLDEN, "stream_%s.npz" % name))
    seed = int(g["rng_seed"])
    random.seed(seed)
    np.random.seed(seed)
    v = mods.video.Video(_Grabber(), ticks_per_second=14700., mode=mods.video_mode.VideoMode.HGR,
                         palette=mods.palette.Palette.NTSC)
    got = []
    with contextlib.redirect_stdout(io.StringIO()):
        for frame, is_aux, budget in g["segments"]:
            tgt = mods.screen.HGRBitmap(palette=mods.palette.Palette.NTSC,
                                        main_memory=mods.screen.MemoryMap(1, g["frames"][frame, 0].copy()))
            op_seq = v.encode_frame(tgt, is_aux=False)
            for _ in range(budget):
                page, content, offs = next(op_seq)
                got.append([page, content] + list(offs))
    assert np.array_equal(np.array(got, np.uint8), g["opcodes"])
    assert np.array_equal(v.pixelmap.packed, g["packed"])
    assert np.array_equal(v.update_priority, g["priority_main"])
    assert [random.getrandbits(32) for _ in range(4)] == g["next_python_words"].tolist()
    assert np.random.randint(0, 256, size=4).tolist() == g["next_numpy_bytes"].tolist()


def test_video_sees_foreign_draws_and_host_edits_between_generators(mods, oracle_tables):
    """Between two generators other code may draw from the global generators and edit the
    encoder's arrays (they are plain attributes upstream): the next generator starts from
    exactly that, as the reference's would."""
    from iivision_b200 import synth
    from oracle import scorer
    frames = synth.synthetic_frames("DHGR", 2, 0.6, seed=41)
    segs = synth.movie_schedule("DHGR", 2)
    random.seed(11)
    np.random.seed(11)
    v = mods.video.Video(_Grabber(), 14700., mode=mods.video_mode.VideoMode.DHGR)
    opy, onp = random.Random(11), np.random.RandomState(11)
    ov = scorer.OracleVideo("DHGR", oracle_tables("DHGR"), py_rng=opy, np_rng=onp)
    tgt = otgt = None
    with contextlib.redirect_stdout(io.StringIO()):
        for n, (frame, is_aux, budget) in enumerate(segs):
            if tgt is None or frame != segs[n - 1][0]:
                tgt = mods.screen.DHGRBitmap(
                    palette=mods.palette.Palette.NTSC,
                    main_memory=mods.screen.MemoryMap(1, frames[frame, 0].copy()),
                    aux_memory=mods.screen.MemoryMap(1, frames[frame, 1].copy()))
                otgt = ov.target_bitmap(frames[frame, 0], frames[frame, 1])
            if n == 2:        # somebody else uses the global generators
                assert random.getrandbits(8) == opy.getrandbits(8)
                assert np.random.randint(0, 256, size=3).tolist() == onp.randint(0, 256, size=3).tolist()
            if n == 3:        # python stream only
                random.random(), opy.random()
            if n == 4:        # and the priorities are edited in place
                v.update_priority[3, 10:20] += 7
                ov.update_priority[3, 10:20] += 7
                v.aux_update_priority[5, 0:8] = 0
                ov.aux_update_priority[5, 0:8] = 0
            seq, oseq = v.encode_frame(tgt, bool(is_aux)), ov.encode_frame(otgt, bool(is_aux))
            for _ in range(budget):
                a, b = next(seq), next(oseq)
                assert (a[0], a[1], a[2]) == (int(b[0]), int(b[1]), [int(x) for x in b[2]]), n
            seq.close()      # "between generators": Movie.encode drops the old one first
    assert np.array_equal(v.update_priority, ov.update_priority)
    assert np.array_equal(v.aux_update_priority, ov.aux_update_priority)
    assert np.array_equal(v.pixelmap.packed, ov.pixelmap.packed)
    assert random.getrandbits(32) == opy.getrandbits(32)


def test_static_helpers_against_reference_fixture(mods):
    """_make_header / _make_footer / _body / _fix_column_* / _double_pixels against
    outputs of the reference classes (tests/golden/helpers.npz)."""
    g = np.load(os.path.join(GOLDEN, "helpers.npz"))
    s = mods.screen
    for mode, cls in (("HGR", s.HGRBitmap), ("DHGR", s.DHGRBitmap)):
        words, other = g[mode + "_words"], g[mode + "_other"]
        assert np.array_equal(cls._make_header(words), g[mode + "_header"])
        assert np.array_equal(cls._make_footer(words), g[mode + "_footer"])
        assert cls._make_header(words[0, 0]) == g[mode + "_header"][0, 0]
        fr = g[mode + "_frame"]
        mm = s.MemoryMap(1, fr[0].copy())
        bm = (s.DHGRBitmap(mods.palette.Palette.NTSC, mm, s.MemoryMap(1, fr[1].copy()))
              if mode == "DHGR" else s.HGRBitmap(mods.palette.Palette.NTSC, mm))
        assert np.array_equal(bm._body(), g[mode + "_body"])
        assert np.array_equal(bm._fix_column_left(other, words), g[mode + "_fix_left"])
        assert np.array_equal(bm._fix_column_right(other, words), g[mode + "_fix_right"])
    got = [s.HGRBitmap._double_pixels(v) for v in range(128)]
    assert got == g["double_pixels"].tolist()
    # literal from the reference's own unit test (screen_test.py:489-497)
    assert s.HGRBitmap._double_pixels(0b1010101) == 0b111001100110011
    with pytest.raises(Exception):
        s.DHGRBitmap._part(3, np.uint64(1))


def test_compute_edit_distance_results_do_not_alias(mods, oracle_luts):
    """The pinned staging buffers are recycled only after the caller dropped the array."""
    edp = mods.mdt.compute_substitute_costs(mods.palette.NTSCPalette)
    a = mods.mdt.compute_edit_distance(edp, mods.screen.DHGRBitmap)
    at = 5000 * 8192 + 100                        # row 5000, columns 100..109 (j < i)
    keep = a[0, at:at + 10].copy()
    assert keep.any()
    view = a[1]                                   # a derived view keeps the buffer busy
    edp2 = mods.mdt.compute_substitute_costs(mods.palette.IIGSPalette)
    b = mods.mdt.compute_edit_distance(edp2, mods.screen.DHGRBitmap)
    assert not np.shares_memory(a, b)
    assert not a.flags.writeable and not b.flags.writeable     # documented: read-only result
    assert np.array_equal(a[0, at:at + 10], keep) and not np.array_equal(a, b)
    ptr = a.ctypes.data
    del a, view
    # EditDistanceParams keeps its arrays as CLASS attributes, like the reference
    # (make_data_tables.py:30-52): edp now carries the IIGS costs edp2 was filled with
    c = mods.mdt.compute_edit_distance(edp, mods.screen.DHGRBitmap)
    assert c.ctypes.data == ptr                   # recycled now
    assert np.array_equal(c, b)
    # a recycled buffer is still zero above the diagonal (only j < i is ever copied into it)
    from oracle import tables
    want, _ = tables.build_table("DHGR", oracle_luts[0], triangular=True)
    assert np.array_equal(c, want)


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_emit_stream_device_matches_reference_bytes(mode):
    """iiv_emit_stream against bytes produced by the reference's Movie.emit_stream."""
    import torch
    from iivision_b200 import movie
    g = np.load(os.path.join(GOLDEN, "byte_stream.npz"))
    addresses = (g["tick_addr"], int(g["ack_addr"]), int(g["terminate_addr"]))
    for case in ("short", "frames", "exact", "capped"):
        key = "%s_%s" % (mode.lower(), case)
        rec = torch.from_numpy(g[key + "_records"]).cuda()
        ticks = torch.from_numpy(g[key + "_ticks"]).cuda()
        out = movie.emit_stream_device(mode, rec, ticks, int(g[key + "_max"]) or None, addresses)
        assert np.array_equal(out.cpu().numpy(), g[key + "_bytes"]), key
    empty = movie.emit_stream_device(mode, torch.zeros((0, 8), dtype=torch.uint8, device="cuda"),
                                     torch.zeros((0,), dtype=torch.uint8, device="cuda"),
                                     None, addresses).cpu().numpy()
    assert len(empty) == 2048 and empty[6] == (1 if mode == "DHGR" else 0)
    assert (int(empty[7]) << 8 | int(empty[8])) == int(g["terminate_addr"])
    bad = torch.from_numpy(g["hgr_short_ticks"].copy()).cuda()
    bad[2] = 5
    with pytest.raises(KeyError):
        movie.emit_stream_device(mode, torch.from_numpy(g["hgr_short_records"]).cuda(), bad,
                                 None, addresses)

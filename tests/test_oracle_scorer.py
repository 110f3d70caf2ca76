"""CPU: the numpy scorer/encoder oracle against fixtures produced by the
UNMODIFIED reference (tests/golden/scorer_*.npz, stream_*.npz; generator:
oracle/make_golden.py)."""

import os
import random

import numpy as np
import pytest

from oracle import scorer

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STREAMS = ["dhgr_full", "dhgr_sparse", "hgr_full", "hgr_sparse", "hgr_long_generator", "hgr_exhaust", "dhgr_long_generator"]


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def test_hole_map_matches_layout():
    holes = scorer.SCREEN_HOLES
    assert holes.sum() == 512
    for off in range(256):
        assert holes[:, off].all() == ((off & 127) >= 120)
    assert scorer.xy_to_page_offset(0, 0) == (0, 0)
    assert scorer.xy_to_page_offset(39, 191) == (31, 0xD0 + 39)


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_primitives_against_reference_fixture(oracle_tables, mode):
    g = load("scorer_%s.npz" % mode.lower())
    spec = scorer.SPECS[mode]
    table = oracle_tables(mode)
    fr = g["frames"]
    aux = (lambda k: fr[k, 1].copy()) if mode == "DHGR" else (lambda k: None)
    src = scorer.OracleBitmap(mode, table, fr[0, 0].copy(), aux(0))
    tgt = scorer.OracleBitmap(mode, table, fr[1, 0].copy(), aux(1))
    assert np.array_equal(src.packed, g["src_packed"])
    assert np.array_equal(tgt.packed, g["tgt_packed"])
    for o in range(spec.n_offsets):
        assert np.array_equal(spec.mask_shift(g["words"], o), g["mask_shift"][o])
        for k, v in enumerate(g["values"]):
            assert np.array_equal(spec.masked_update(o, g["words"], np.uint8(v)),
                                  g["masked_update"][o, k])
    for tag in (("main", "aux") if mode == "DHGR" else ("main",)):
        is_aux = tag == "aux"
        dw = tgt.diff_weights(src, is_aux)
        assert np.array_equal(dw, g["diff_weights_" + tag])
        for (page, content), want in zip(g["delta_cases_" + tag], g["delta_" + tag]):
            got = tgt.compute_delta_page(int(page), np.uint8(content), dw[page, :], is_aux)
            assert np.array_equal(got, want)
        for bo, page, off, content, want in g["pair_difference_" + tag]:
            got = tgt.byte_pair_difference(int(bo), tgt.packed[page, off // 2], np.uint8(content))
            assert int(got) == want
    for page, off, is_aux, val in g["apply_stores"]:
        src.apply(int(page), int(off), bool(is_aux), np.uint8(val))
    assert np.array_equal(src.packed, g["apply_packed"])
    assert np.array_equal(src.main, g["apply_main"])
    if mode == "DHGR":
        assert np.array_equal(src.aux, g["apply_aux"])
    # incremental fix-ups stay consistent with a full repack
    want = src.packed.copy()
    src.repack()
    assert np.array_equal(src.packed, want)


@pytest.mark.parametrize("name", STREAMS)
def test_opcode_stream_against_reference_fixture(oracle_tables, name):
    g = load("stream_%s.npz" % name)
    mode = str(g["mode"])
    seed = int(g["rng_seed"])
    py, npr = random.Random(seed), np.random.RandomState(seed)
    v = scorer.OracleVideo(mode, oracle_tables(mode), py_rng=py, np_rng=npr)
    frames = g["frames"]
    ops, real, sims = [], [], []
    for frame, is_aux, budget in g["segments"]:
        tgt = v.target_bitmap(frames[frame, 0], frames[frame, 1] if mode == "DHGR" else None)
        v.out_of_work = {True: False, False: False}
        seq = v.encode_frame(tgt, bool(is_aux))
        for _ in range(budget):
            page, content, offs = next(seq)
            real.append(0 if v.out_of_work[bool(is_aux)] else 1)
            ops.append([page, content] + list(offs))
        sims.append(v.mean_priority)
    assert np.array_equal(np.array(ops, np.uint8), g["opcodes"])
    assert np.array_equal(np.array(real, np.uint8), g["real"])
    assert np.allclose(sims, g["similarity"], rtol=0, atol=0)
    assert np.array_equal(v.pixelmap.packed, g["packed"])
    assert np.array_equal(v.main, g["main"])
    assert np.array_equal(v.update_priority, g["priority_main"])
    if mode == "DHGR":
        assert np.array_equal(v.aux, g["aux"])
        assert np.array_equal(v.aux_update_priority, g["priority_aux"])
    # both MT19937 streams were consumed exactly as the reference consumed them
    assert [py.getrandbits(32) for _ in range(4)] == g["next_python_words"].tolist()
    assert npr.randint(0, 256, size=4).tolist() == g["next_numpy_bytes"].tolist()


def test_reference_unit_test_literals(oracle_tables):
    """video_test.py:28-43, 48-79: packed words and the pair-index arithmetic."""
    table = oracle_tables("DHGR")
    zeros = lambda: np.zeros((32, 256), np.uint8)
    src = scorer.OracleBitmap("DHGR", table, zeros(), zeros())
    aux = zeros()
    aux[0, 0], aux[0, 1] = 0b1111111, 0b1010101
    tgt = scorer.OracleBitmap("DHGR", table, zeros(), aux)
    assert tgt.packed[0, 0] == 0b0000000000101010100000001111111000
    diff = tgt.diff_weights(src, True)
    assert diff[0, 0] == table[0][0b0001111111000]
    assert diff[0, 1] == table[2][0b0001010101000]
    src = scorer.OracleBitmap("DHGR", table, zeros(), aux.copy())
    aux2 = zeros()
    aux2[0, 0], aux2[0, 1] = 0b1101101, 0b0110110
    tgt = scorer.OracleBitmap("DHGR", table, zeros(), aux2)
    assert tgt.packed[0, 0] == 0b0000000000011011000000001101101000
    diff = tgt.diff_weights(src, True)
    assert diff[0, 0] == table[0][0b00011111110000001101101000]
    assert diff[0, 1] == table[2][0b00010101010000000110110000]

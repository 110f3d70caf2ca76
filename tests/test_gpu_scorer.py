"""Parity of the CUDA scoring primitives against the numpy oracle (bit-exact)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from iivision_b200 import ops
    return ops


def _torch():
    import torch
    return torch


def _frames(mode, n, f, seed):
    from iivision_b200.synth import synthetic_frames
    return synthetic_frames(mode, n, f, seed)


def _dev(a):
    return _torch().from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_pack(ops, mode):
    from oracle import scorer
    spec = scorer.SPECS[mode]
    fr = _frames(mode, 5, 1.0, 3)
    # include holes with data and extreme bytes: _pack itself does not care
    fr[0][:] = 0xFF
    fr[1][:] = np.random.default_rng(1).integers(0, 256, fr[1].shape, dtype=np.uint8)
    main = _dev(fr[:, 0])
    aux = _dev(fr[:, 1]) if mode == "DHGR" else None
    got = ops.pack(mode, main, aux).cpu().numpy().view(np.uint64)
    for k in range(fr.shape[0]):
        want = spec.pack(fr[k, 0], fr[k, 1] if mode == "DHGR" else None)
        assert np.array_equal(got[k], want)


def test_pack_reference_literals(ops):
    """Golden words from the reference's own tests (video_test.py:28-30, 63-66;
    screen_test.py DHGR/HGR pack cases)."""
    aux = np.zeros((1, 32, 256), np.uint8)
    main = np.zeros((1, 32, 256), np.uint8)
    aux[0, 0, 0] = 0b1111111
    aux[0, 0, 1] = 0b1010101
    got = ops.pack("DHGR", _dev(main), _dev(aux)).cpu().numpy().view(np.uint64)
    assert got[0, 0, 0] == 0b0000000000101010100000001111111000
    aux[0, 0, 0] = 0b1101101
    aux[0, 0, 1] = 0b0110110
    got = ops.pack("DHGR", _dev(main), _dev(aux)).cpu().numpy().view(np.uint64)
    assert got[0, 0, 0] == 0b0000000000011011000000001101101000


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_elementwise_ops(ops, mode):
    from oracle import scorer
    spec = scorer.SPECS[mode]
    rng = np.random.default_rng(5)
    width = spec.header_bits + spec.body_bits + spec.footer_bits
    words = rng.integers(0, 1 << width, size=(7, 128), dtype=np.uint64)
    d = _dev(words.view(np.int64))
    for o in range(spec.n_offsets):
        got = ops.mask_and_shift(mode, o, d).cpu().numpy().view(np.uint64)
        assert np.array_equal(got, spec.mask_shift(words, o))
        for v in (0, 1, 0x55, 0x7F, 0x80, 0xFF):
            got = ops.masked_update(mode, o, d, v).cpu().numpy().view(np.uint64)
            assert np.array_equal(got, spec.masked_update(o, words, np.uint8(v)))
        bm = scorer.OracleBitmap(mode, None, np.zeros((32, 256), np.uint8),
                                 np.zeros((32, 256), np.uint8))
        want = words.copy()
        bm._fix_array_neighbours(want, o)
        got = ops.fix_array_neighbours(mode, o, d.clone()).cpu().numpy().view(np.uint64)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_diff_weights_and_delta(ops, oracle_tables, device_tables, mode):
    from oracle import scorer
    tab = oracle_tables(mode)
    dtab = device_tables(mode)
    fr = _frames(mode, 4, 0.6, 11)
    banks = (False, True) if mode == "DHGR" else (False,)
    src = scorer.OracleBitmap(mode, tab, fr[0, 0], fr[0, 1] if mode == "DHGR" else None)
    tgt = scorer.OracleBitmap(mode, tab, fr[3, 0], fr[3, 1] if mode == "DHGR" else None)
    dsrc = _dev(src.packed.view(np.int64))
    dtgt = _dev(tgt.packed.view(np.int64))
    for is_aux in banks:
        want = tgt.diff_weights(src, is_aux)
        got = ops.diff_weights(mode, is_aux, dsrc, dtgt, dtab).cpu().numpy()
        assert np.array_equal(got, want)
        for content in (0, 0x2A, 0x7F) + ((0x80, 0xFF) if mode == "HGR" else ()):
            want_c = tgt.diff_weights(src, is_aux, content=np.uint8(content))
            got_c = ops.diff_weights(mode, is_aux, dsrc, dtgt, dtab, content).cpu().numpy()
            assert np.array_equal(got_c, want_c)
            for page in (0, 13, 31):
                wp = tgt.compute_delta_page(page, np.uint8(content), want[page], is_aux)
                gp = ops.compute_delta_page(mode, is_aux, dtgt, page, content,
                                            _dev(want[page]), dtab).cpu().numpy()
                assert np.array_equal(gp, wp)
                wpg = tgt.diff_weights_page(src.packed[page], tgt.packed[page], is_aux,
                                            np.uint8(content))
                gpg = ops.diff_weights_page(mode, is_aux, dsrc[page], dtgt[page], dtab,
                                            content).cpu().numpy()
                assert np.array_equal(gpg, wpg)


def test_diff_weights_reference_index_goldens(ops, device_tables):
    """video_test.py:36-43, 70-79: pins the pair-index arithmetic."""
    dtab = device_tables("DHGR")
    t = dtab.cpu().numpy()
    main = np.zeros((32, 256), np.uint8)
    aux = np.zeros((32, 256), np.uint8)
    aux[0, 0], aux[0, 1] = 0b1111111, 0b1010101
    zero = ops.pack("DHGR", _dev(main), _dev(main))
    tgt = ops.pack("DHGR", _dev(main), _dev(aux))
    diff = ops.diff_weights("DHGR", True, zero, tgt, dtab).cpu().numpy()
    assert diff[0, 0] == t[0][0b0001111111000]
    assert diff[0, 1] == t[2][0b0001010101000]
    aux2 = np.zeros((32, 256), np.uint8)
    aux2[0, 0], aux2[0, 1] = 0b1101101, 0b0110110
    tgt2 = ops.pack("DHGR", _dev(main), _dev(aux2))
    diff = ops.diff_weights("DHGR", True, tgt, tgt2, dtab).cpu().numpy()
    assert diff[0, 0] == t[0][0b00011111110000001101101000]
    assert diff[0, 1] == t[2][0b00010101010000000110110000]


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_delta_rows(ops, oracle_tables, device_tables, mode):
    from oracle import scorer
    tab, dtab = oracle_tables(mode), device_tables(mode)
    fr = _frames(mode, 2, 1.0, 21)
    tgt = scorer.OracleBitmap(mode, tab, fr[1, 0], fr[1, 1] if mode == "DHGR" else None)
    dtgt = _dev(tgt.packed.view(np.int64))
    zero_row = np.zeros(256, np.int32)
    for is_aux in ((False, True) if mode == "DHGR" else (False,)):
        got = ops.delta_rows(mode, is_aux, dtgt, dtab).cpu().numpy()
        assert got.shape == (32, 256 if mode == "HGR" else 128, 256)
        rng = np.random.default_rng(2)
        for _ in range(40):
            page = int(rng.integers(32))
            content = int(rng.integers(got.shape[1]))
            want = tgt.compute_delta_page(page, np.uint8(content), zero_row, is_aux)
            assert np.array_equal(got[page, content].astype(np.int32), want)


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_apply_and_pair_difference(ops, oracle_tables, device_tables, mode):
    from oracle import scorer
    torch = _torch()
    tab, dtab = oracle_tables(mode), device_tables(mode)
    spec = scorer.SPECS[mode]
    fr = _frames(mode, 1, 1.0, 31)
    main = fr[0, 0].copy()
    aux = fr[0, 1].copy() if mode == "DHGR" else None
    bm = scorer.OracleBitmap(mode, tab, main, aux)
    dmain = _dev(main)
    daux = _dev(aux) if aux is not None else None
    dpacked = ops.pack(mode, dmain, daux)
    rng = np.random.default_rng(9)
    stores = []
    for k in range(300):
        page = int(rng.integers(32))
        # exercise row ends (no leak across page ends) as the reference's tests do
        offset = int(rng.choice([0, 1, 2, 127, 128, 253, 254, 255, int(rng.integers(256))]))
        is_aux = bool(rng.integers(2)) if mode == "DHGR" else False
        value = int(rng.integers(256 if mode == "HGR" else 128))
        stores.append((page, offset, int(is_aux), value))
        bm.apply(page, offset, is_aux, np.uint8(value))
    ops.apply_stores(mode, dpacked, dmain, daux, stores)
    assert np.array_equal(dpacked.cpu().numpy().view(np.uint64), bm.packed)
    assert np.array_equal(dmain.cpu().numpy(), bm.main)
    if aux is not None:
        assert np.array_equal(daux.cpu().numpy(), bm.aux)
    # incremental fix-ups are consistent with a full repack
    assert torch.equal(ops.pack(mode, dmain, daux), dpacked)

    words = bm.packed.reshape(-1)[:512]
    contents = rng.integers(0, 256 if mode == "HGR" else 128, size=512).astype(np.uint8)
    for o in range(spec.n_offsets):
        got = ops.byte_pair_difference(mode, o, _dev(words.view(np.int64)),
                                       _dev(contents), dtab).cpu().numpy()
        want = [bm.byte_pair_difference(o, words[k], contents[k]) for k in range(512)]
        assert got.tolist() == [int(x) for x in want]

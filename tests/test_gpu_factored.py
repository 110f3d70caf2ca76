"""The scorer without the table: iiv_score_frames_factored (edit-distance entries evaluated
from per-segment factor tables in shared memory) against iiv_score_frames (entries gathered
from the table in HBM) and, through it, the oracle -- bit-exact."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(mode, n, fraction, seed):
    import torch
    from iivision_b200 import ops, synth
    fr = synth.synthetic_frames(mode, n + 1, fraction, seed=seed)
    banks = fr.shape[1]
    d = torch.from_numpy(fr).cuda()
    packed = ops.pack(mode, d[:, 0].contiguous(), d[:, 1].contiguous() if banks == 2 else None)
    return packed[:n].contiguous(), d[1:].contiguous(), packed[1:].contiguous(), banks


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
@pytest.mark.parametrize("pid", [5, 0])
def test_factored_equals_table_path(mode, pid, oracle_luts, device_tables):
    import torch
    from iivision_b200 import ops
    table = device_tables(mode, pid)
    factors = ops.score_factors(mode, oracle_luts[pid])
    rng = np.random.default_rng(3)
    for n, fraction, seed in ((1, 1.0, 1), (7, 1.0, 2), (40, 0.3, 3), (333, 0.05, 4)):
        src, tgt_mem, tgt_packed, banks = _frames(mode, n, fraction, seed)
        prio0 = rng.integers(0, 5000, size=(n, banks, 32, 256)).astype(np.int32)
        prio0[rng.random(prio0.shape) < 0.3] = 0
        pa, pb = torch.from_numpy(prio0.copy()).cuda(), torch.from_numpy(prio0.copy()).cuda()
        tp_a, diff_a = ops.score_frames(mode, src, tgt_mem, table, priority=pa)
        tp_b, diff_b = ops.score_frames(mode, src, tgt_mem, factors=factors, priority=pb)
        assert torch.equal(tp_b, tgt_packed) and torch.equal(tp_a, tp_b)
        assert torch.equal(diff_a, diff_b)
        assert torch.equal(pa, pb)
    # one source bitmap for every frame, holes kept, no priorities, no packed output
    src, tgt_mem, _, banks = _frames(mode, 5, 1.0, 9)
    _, da = ops.score_frames(mode, src[0], tgt_mem, table, zero_holes=False, want_packed=False)
    _, db = ops.score_frames(mode, src[0], tgt_mem, factors=factors, zero_holes=False,
                             want_packed=False)
    assert torch.equal(da, db)


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_factored_extreme_luts(mode):
    """Packed 16-bit arithmetic at its limits: every substitution 255 (entries up to
    n * 255; the INF of an absent swap must never win), all-zero costs, and an adversarial
    random LUT with zeros off the diagonal."""
    import torch
    from iivision_b200 import ops
    rng = np.random.default_rng(11)
    adv = rng.integers(0, 256, size=(16, 16)).astype(np.int32)
    adv[rng.random((16, 16)) < 0.15] = 0
    adv = np.minimum(adv, adv.T)
    np.fill_diagonal(adv, 0)
    hi = np.full((16, 16), 255, dtype=np.int32)
    src, tgt_mem, _, banks = _frames(mode, 24, 1.0, 21)
    for lut in (adv, hi, np.zeros((16, 16), dtype=np.int32)):
        table = ops.table_generate(mode, lut)
        factors = ops.score_factors(mode, lut)
        _, da = ops.score_frames(mode, src, tgt_mem, table)
        _, db = ops.score_frames(mode, src, tgt_mem, factors=factors)
        assert torch.equal(da, db)
        del table


def test_factored_bad_arguments(oracle_luts):
    from iivision_b200 import ops
    from iivision_b200._lib import IIVError
    lut = oracle_luts[5].copy()
    lut[2, 3] = 256
    with pytest.raises(IIVError):
        ops.score_factors("DHGR", lut)
    src, tgt_mem, _, _ = _frames("DHGR", 2, 1.0, 1)
    with pytest.raises(ValueError):
        ops.score_frames("DHGR", src, tgt_mem)


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_asymmetric_luts_keep_their_orientation(mode):
    """The C ABI takes any 16x16 costs in 0..255, symmetric or not (the reference's are
    symmetric): S[a][b] with a from the source's string, b from the target's, in the chain
    kernel, the split generator and the factored scorer alike."""
    import torch
    from iivision_b200 import ops
    from iivision_b200._lib import ALGO_CHAIN, ALGO_SPLIT
    rng = np.random.default_rng(23)
    src, tgt_mem, _, _ = _frames(mode, 16, 1.0, 31)
    for _ in range(2):
        lut = rng.integers(0, 256, size=(16, 16)).astype(np.int32)
        lut[rng.random((16, 16)) < 0.2] = 0
        a = ops.table_generate(mode, lut, algo=ALGO_CHAIN)
        b = ops.table_generate(mode, lut, algo=ALGO_SPLIT)
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))
        del a
        _, da = ops.score_frames(mode, src, tgt_mem, b)
        _, db = ops.score_frames(mode, src, tgt_mem, factors=ops.score_factors(mode, lut))
        assert torch.equal(da, db)
        del b

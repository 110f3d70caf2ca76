"""GPU: randomised differential test of the encoder pipeline against the numpy oracle.
Covers what the fixed cases do not reach on purpose: tiny and large budgets, schedules
that flip banks at odd places, frames with long flat runs (many equal deltas -> contender
overflow and page conflicts), nearly static frames (heap runs dry, re-queued cells)."""

import numpy as np
import pytest

from encoder_util import run_device, run_oracle

pytestmark = pytest.mark.gpu


def _flat_frames(mode, n_frames, rng):
    """Frames made of long runs of one byte value with a few changes per frame."""
    from iivision_b200.synth import screen_holes
    holes = screen_holes()
    hi = 256 if mode == "HGR" else 128
    banks = 2 if mode == "DHGR" else 1
    out = np.zeros((n_frames, banks, 32, 256), np.uint8)
    cur = np.zeros((banks, 32, 256), np.uint8)
    for k in range(n_frames):
        for b in range(banks):
            for _ in range(int(rng.integers(1, 6))):
                page = int(rng.integers(0, 32))
                lo = int(rng.integers(0, 200))
                cur[b, page, lo:lo + int(rng.integers(8, 56))] = rng.integers(0, hi)
            cur[b][holes] = 0
        out[k] = cur
    return out


@pytest.mark.parametrize("case", range(int(__import__("os").environ.get("IIV_RANDOM_CASES", "120"))))
def test_random_configurations(oracle_tables, device_tables, case):
    from iivision_b200 import ops
    from iivision_b200.synth import synthetic_frames
    rng = np.random.default_rng(1000 + case)
    mode = "DHGR" if case % 3 else "HGR"
    n_frames = int(rng.integers(1, 4))
    if case % 2:
        frames = _flat_frames(mode, n_frames, rng)
    else:
        frames = synthetic_frames(mode, n_frames, float(rng.choice([1.0, 0.5, 0.1, 0.02])),
                                  seed=int(rng.integers(0, 1 << 30)))
    segs = []
    aux = 0
    for fr in range(n_frames):
        for _ in range(int(rng.integers(1, 5))):
            budget = int(rng.choice([0, 1, 2, 7, 33, 150, 292, 700]))
            segs.append((fr, aux if mode == "DHGR" else 0, budget))
            if mode == "DHGR" and rng.random() < 0.7:
                aux ^= 1
    seed = int(rng.integers(0, 1 << 20))
    want_ops, want_real, v, py, npr = run_oracle(mode, oracle_tables(mode), frames, segs, seed)
    got, info, states = run_device(ops, mode, device_tables(mode), frames[None], segs, [seed])
    got = got[0]
    assert got.shape[0] == want_ops.shape[0]
    if want_ops.shape[0]:
        mism = np.flatnonzero((got[:, :6].astype(np.int64) != want_ops).any(axis=1))
        assert mism.size == 0, "case %d (%s): first differing opcode %d of %d: got %s want %s" % (
            case, mode, mism[0], len(got), got[mism[0]], want_ops[mism[0]])
        assert np.array_equal(got[:, 6], want_real)
    import torch
    packed = ops.state_field(states, ops.F_PACKED, torch.int64, (32, 128))[0]
    assert np.array_equal(packed.cpu().numpy().view(np.uint64), v.pixelmap.packed)
    pm = ops.state_field(states, ops.F_PRIO_MAIN, torch.int32, (32, 256))[0]
    assert np.array_equal(pm.cpu().numpy(), v.update_priority)
    mt_py = ops.state_field(states, ops.F_MT_PY, torch.int32, (640,))[0]
    assert ops.mt_to_python(mt_py.cpu().numpy().view(np.uint32)[:625])[1] == py.getstate()[1]


@pytest.mark.parametrize("mode,fraction", [("DHGR", 1.0), ("DHGR", 0.1), ("HGR", 1.0)])
def test_large_budgets_between_small_ones(oracle_tables, device_tables, mode, fraction):
    """Budgets above a third of the heap switch the prefix selection off (the whole heap is
    sorted and fills the key array, so nothing can be prepared for the next segment behind
    it); smaller ones before and after switch it back on."""
    from iivision_b200 import ops
    from iivision_b200.synth import synthetic_frames
    frames = synthetic_frames(mode, 3, fraction, seed=77)
    a = 1 if mode == "DHGR" else 0
    segs = [(0, 0, 292), (0, a, 1400), (1, 0, 2048), (1, a, 292), (2, 0, 5), (2, a, 1366),
            (2, 0, 1365)]
    want_ops, want_real, v, py, npr = run_oracle(mode, oracle_tables(mode), frames, segs, 11)
    got, info, states = run_device(ops, mode, device_tables(mode), frames[None], segs, [11])
    assert np.array_equal(got[0][:, :6].astype(np.int64), want_ops)
    assert np.array_equal(got[0][:, 6], want_real)
    import torch
    mt_np = ops.state_field(states, ops.F_MT_NP, torch.int32, (640,))[0]
    st = npr.get_state()
    words = mt_np.cpu().numpy().view(np.uint32)
    # same position in stream N: the next draws agree
    from iivision_b200.ops import mt_to_numpy
    r = np.random.RandomState()
    r.set_state(mt_to_numpy(words[:625]))
    assert np.array_equal(r.randint(0, 256, 8), npr.randint(0, 256, 8)), st[2]

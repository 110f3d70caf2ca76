"""Bit-exact opcode streams: CUDA encoder vs the numpy oracle of video.py."""

import numpy as np
import pytest

from encoder_util import run_device, run_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from iivision_b200 import ops
    return ops


def _check_clip(ops, mode, tab, dtab, frames, segments, seed):
    import torch
    want_ops, want_real, v, py, npr = run_oracle(mode, tab, frames, segments, seed)
    got, info, states = run_device(ops, mode, dtab, frames[None], segments, [seed])
    got = got[0]
    assert got.shape[0] == want_ops.shape[0]
    mism = np.flatnonzero((got[:, :6].astype(np.int64) != want_ops).any(axis=1))
    assert mism.size == 0, "first differing opcode %d: got %s want %s" % (
        mism[0], got[mism[0]], want_ops[mism[0]])
    assert np.array_equal(got[:, 6], want_real)
    # encoder state after the run
    packed = ops.state_field(states, ops.F_PACKED, torch.int64, (32, 128))[0]
    assert np.array_equal(packed.cpu().numpy().view(np.uint64), v.pixelmap.packed)
    main = ops.state_field(states, ops.F_MAIN, torch.uint8, (32, 256))[0]
    assert np.array_equal(main.cpu().numpy(), v.main)
    pm = ops.state_field(states, ops.F_PRIO_MAIN, torch.int32, (32, 256))[0]
    assert np.array_equal(pm.cpu().numpy(), v.update_priority)
    if mode == "DHGR":
        aux = ops.state_field(states, ops.F_AUX, torch.uint8, (32, 256))[0]
        assert np.array_equal(aux.cpu().numpy(), v.aux)
        pa = ops.state_field(states, ops.F_PRIO_AUX, torch.int32, (32, 256))[0]
        assert np.array_equal(pa.cpu().numpy(), v.aux_update_priority)
    # both MT19937 streams end in exactly the reference's state
    mt_py = ops.state_field(states, ops.F_MT_PY, torch.int32, (640,))[0]
    mt_py = mt_py.cpu().numpy().view(np.uint32)[:625]
    assert ops.mt_to_python(mt_py)[1] == py.getstate()[1]
    mt_np = ops.state_field(states, ops.F_MT_NP, torch.int32, (640,))[0]
    mt_np = mt_np.cpu().numpy().view(np.uint32)[:625]
    st = npr.get_state()
    assert np.array_equal(mt_np[:624], st[1]) and int(mt_np[624]) == st[2]
    return info[0]


@pytest.mark.parametrize("mode,fraction", [
    ("DHGR", 1.0), ("DHGR", 0.05), ("HGR", 1.0), ("HGR", 0.05)])
def test_movie_schedule_stream(ops, oracle_tables, device_tables, mode, fraction):
    from iivision_b200.synth import movie_schedule, synthetic_frames
    frames = synthetic_frames(mode, 4, fraction, seed=1)
    segs = movie_schedule(mode, 4)
    _check_clip(ops, mode, oracle_tables(mode), device_tables(mode), frames, segs, 0)


def test_out_of_work_and_requeue(ops, oracle_tables, device_tables):
    """Tiny change set: the first-pass heap runs dry, re-queued cells are popped
    and finally the generator pads with (32, target[0,0], [0,0,0,0])."""
    from iivision_b200.synth import synthetic_frames
    mode = "DHGR"
    frames = synthetic_frames(mode, 3, 0.01, seed=4)
    segs = [(0, 0, 300), (0, 1, 300), (1, 0, 0), (1, 1, 200), (1, 0, 200),
            (2, 0, 150), (2, 1, 150), (2, 0, 1)]
    info = _check_clip(ops, mode, oracle_tables(mode), device_tables(mode),
                       frames, segs, 3)
    assert (info[:, 0] < np.array([s[2] for s in segs])).any()


def test_identical_target_emits_only_padding(ops, oracle_tables, device_tables):
    from iivision_b200.synth import synthetic_frames
    frames = np.zeros((1, 1, 32, 256), np.uint8)
    info = _check_clip(ops, "HGR", oracle_tables("HGR"), device_tables("HGR"),
                       frames, [(0, 0, 10)], 5)
    assert info[0, 0] == 0


def test_batched_clips_independent(ops, oracle_tables, device_tables):
    """Several clips in one launch equal the same clips run one by one."""
    from iivision_b200.synth import movie_schedule, synthetic_frames
    mode = "DHGR"
    clips = np.stack([synthetic_frames(mode, 2, f, seed=s)
                      for f, s in ((1.0, 10), (0.3, 11), (0.05, 12), (1.0, 13), (0.0, 14))])
    segs = movie_schedule(mode, 2, opcodes_per_frame=400, flip_every=150)
    seeds = [100, 101, 102, 103, 104]
    got, info, _ = run_device(ops, mode, device_tables(mode), clips, segs, seeds)
    for k in range(clips.shape[0]):
        want_ops, want_real, *_ = run_oracle(mode, oracle_tables(mode), clips[k], segs, seeds[k])
        assert np.array_equal(got[k][:, :6].astype(np.int64), want_ops)
        assert np.array_equal(got[k][:, 6], want_real)


def test_mt19937_device_streams(ops):
    """Device MT19937 equals CPython random / numpy legacy word for word."""
    import random
    import torch
    for seed, n in ((0, 1), (1, 623), (2, 624), (3, 625), (4, 5000)):
        r = random.Random(seed)
        mt = torch.from_numpy(
            np.pad(ops.mt_from_python(r.getstate()), (0, 15)).view(np.int32)).cuda()
        words = ops.mt_draw(mt, n).cpu().numpy().view(np.uint32)
        want = np.array([r.getrandbits(32) for _ in range(n)], dtype=np.uint32)
        assert np.array_equal(words, want)
        assert ops.mt_to_python(mt.cpu().numpy().view(np.uint32)[:625])[1] == r.getstate()[1]
        rs = np.random.RandomState(seed)
        mt = torch.from_numpy(
            np.pad(ops.mt_from_numpy(rs.get_state()), (0, 15)).view(np.int32)).cuda()
        words = ops.mt_draw(mt, n).cpu().numpy().view(np.uint32)
        assert np.array_equal(words & 0xFF, rs.randint(0, 256, size=n).astype(np.uint32))


def test_bad_schedule_rejected(ops, device_tables):
    import torch
    from iivision_b200._lib import IIVError
    states = ops.new_clip_states(1)
    tmem = torch.zeros((1, 1, 1, 32, 256), dtype=torch.uint8, device="cuda")
    tp = torch.zeros((1, 1, 32, 128), dtype=torch.int64, device="cuda")
    for segs in ([(1, 0, 5)], [(0, 1, 5)], [(0, 0, (1 << 17) + 1)]):
        with pytest.raises(IIVError):
            ops.encode_clips("HGR", states, tmem, tp, segs, device_tables("HGR"))


def test_planned_schedule_equals_plain_call(ops, device_tables):
    """iiv_encode_clips_planned (schedule resident on the device) = iiv_encode_clips."""
    import torch
    from encoder_util import seed_states
    from iivision_b200.synth import movie_schedule, synthetic_frames
    mode = "DHGR"
    frames = synthetic_frames(mode, 2, 0.6, seed=21)[None]
    segs = movie_schedule(mode, 2, opcodes_per_frame=300, flip_every=110)
    tmem = torch.from_numpy(frames).cuda()
    tpacked = ops.pack(mode, tmem[0, :, 0].contiguous(), tmem[0, :, 1].contiguous()).view(1, 2, 32, 128)
    outs = []
    for plan in (segs, ops.SegmentPlan(segs)):
        st = ops.new_clip_states(1)
        seed_states(ops, st, [9])
        opc, info = ops.encode_clips(mode, st, tmem, tpacked, plan, device_tables(mode))
        torch.cuda.synchronize()
        outs.append((opc.cpu().numpy(), info.cpu().numpy()[..., :4], st.cpu().numpy()))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)

"""BASELINE.json configs[3] and configs[4] against the UNMODIFIED reference.

tests/golden/long_streams.json holds SHA-256 digests of the opcode streams the reference's
Video.encode_frame (video.py:72-301) emits for a 600-frame prefix of the 6000-frame DHGR clip
and for 8 of the 64 batch clips (oracle/make_golden.py --long-only).  CPU: the numpy oracle
reproduces the first checkpoints.  GPU: the CUDA encoder reproduces all of them, and a full
chip of distinct clips matches the oracle.
"""

import hashlib
import json
import os

import numpy as np
import pytest

from encoder_util import run_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "long_streams.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def test_workloads_are_the_ones_hashed(golden):
    from iivision_b200 import synth
    assert golden["long_clip"] == synth.LONG_CLIP
    bc = dict(synth.BATCH_CLIPS)
    bc["golden_clips"] = list(bc["golden_clips"])
    assert golden["batch_clips"] == bc
    assert sorted(int(k) for k in golden["batch"]) == sorted(synth.BATCH_CLIPS["golden_clips"])


def test_oracle_reproduces_reference_digests(golden, oracle_tables):
    """oracle/scorer.py against the reference's own stream on the named workloads (first
    3 frames of the long clip, first frame of two batch clips)."""
    from iivision_b200 import synth
    tab = oracle_tables("DHGR")
    frames = synth.long_clip_frames(3)
    ops, *_ = run_oracle("DHGR", tab, frames, synth.movie_schedule("DHGR", 3),
                         synth.LONG_CLIP["rng_seed"])
    assert synth.opcode_digest(ops[:980]) == golden["long"]["opcode_sha256"]["1"]
    assert synth.opcode_digest(ops) == golden["long"]["opcode_sha256"]["3"]
    for clip in (0, 63):
        fr = synth.batch_clip_frames(clip, 1)
        ops, *_ = run_oracle("DHGR", tab, fr, synth.movie_schedule("DHGR", 1),
                             synth.batch_clip_seeds(clip)[1])
        assert synth.opcode_digest(ops) == golden["batch"][str(clip)]["opcode_sha256"]["1"]


def _state_digest(ops_mod, states, k=0):
    import torch
    h = hashlib.sha256()
    for field, dt, shape in ((ops_mod.F_PACKED, torch.int64, (32, 128)),
                             (ops_mod.F_MAIN, torch.uint8, (32, 256)),
                             (ops_mod.F_AUX, torch.uint8, (32, 256)),
                             (ops_mod.F_PRIO_MAIN, torch.int32, (32, 256)),
                             (ops_mod.F_PRIO_AUX, torch.int32, (32, 256))):
        h.update(ops_mod.state_field(states, field, dt, shape)[k].cpu().numpy().tobytes())
    return h.hexdigest()


def _next_words(ops_mod, states, k=0):
    import torch
    mt_py = ops_mod.state_field(states, ops_mod.F_MT_PY, torch.int32, (640,))[k].clone()
    mt_np = ops_mod.state_field(states, ops_mod.F_MT_NP, torch.int32, (640,))[k].clone()
    py = int(ops_mod.mt_draw(mt_py, 1).cpu().numpy().view(np.uint32)[0])
    npb = int(ops_mod.mt_draw(mt_np, 1).cpu().numpy().view(np.uint32)[0]) & 0xFF
    return py, npb


@pytest.mark.gpu
def test_long_clip_prefix_matches_reference(golden, device_tables):
    """configs[3]: 600 frames = 588 000 opcodes in one launch, digests at every checkpoint,
    final encoder state and both RNG positions."""
    import torch
    from iivision_b200 import ops, synth
    g = golden["long"]
    n = g["n_frames"]
    frames = synth.long_clip_frames(n)
    plan = ops.SegmentPlan(synth.movie_schedule("DHGR", n))
    tmem = torch.from_numpy(frames[None]).cuda()
    tpacked = ops.pack("DHGR", tmem[0, :, 0].contiguous(),
                       tmem[0, :, 1].contiguous()).view(1, n, 32, 128)
    states = ops.seed_clip_states(ops.new_clip_states(1), [synth.LONG_CLIP["rng_seed"]])
    opcodes, _ = ops.encode_clips("DHGR", states, tmem, tpacked, plan, device_tables("DHGR"))
    torch.cuda.synchronize()
    got = opcodes.cpu().numpy()[0]
    for frames_done, want in sorted(g["opcode_sha256"].items(), key=lambda kv: int(kv[0])):
        assert synth.opcode_digest(got[:int(frames_done) * 980]) == want, frames_done
    assert _state_digest(ops, states) == g["state_sha256"]
    assert _next_words(ops, states) == (g["next_python_word"], g["next_numpy_byte"])


@pytest.mark.gpu
def test_batch_clips_match_reference(golden, device_tables):
    """configs[4]: the 64 clips in one launch; the 8 whose reference digests are committed
    are compared (stream, state, RNG positions)."""
    import torch
    from iivision_b200 import ops, synth
    bc = synth.BATCH_CLIPS
    n_clips, n = bc["n_clips"], bc["n_frames"]
    clips = np.stack([synth.batch_clip_frames(c) for c in range(n_clips)])
    plan = ops.SegmentPlan(synth.movie_schedule("DHGR", n))
    tmem = torch.from_numpy(clips).cuda()
    flat = tmem.view(-1, 2, 32, 256)
    tpacked = ops.pack("DHGR", flat[:, 0].contiguous(), flat[:, 1].contiguous()).view(
        n_clips, n, 32, 128)
    states = ops.seed_clip_states(ops.new_clip_states(n_clips),
                                  [synth.batch_clip_seeds(c)[1] for c in range(n_clips)])
    opcodes, _ = ops.encode_clips("DHGR", states, tmem, tpacked, plan, device_tables("DHGR"))
    torch.cuda.synchronize()
    got = opcodes.cpu().numpy()
    for key, g in golden["batch"].items():
        c = int(key)
        for frames_done, want in g["opcode_sha256"].items():
            assert synth.opcode_digest(got[c][:int(frames_done) * 980]) == want, (c, frames_done)
        assert _state_digest(ops, states, c) == g["state_sha256"], c
        assert _next_words(ops, states, c) == (g["next_python_word"], g["next_numpy_byte"]), c


@pytest.mark.gpu
def test_full_chip_of_distinct_clips_matches_oracle(oracle_tables, device_tables):
    """148 DISTINCT clips (frames and seeds) x 2 frames, one block per SM: every clip's
    stream equals the oracle's.  The lock-free opcode pipeline is timing-sensitive by
    construction; distinct clips give every block its own L2 latencies and interleavings."""
    from iivision_b200 import ops, synth
    from encoder_util import run_device
    n_clips, n = 148, 2
    tab = oracle_tables("DHGR")
    segs = synth.movie_schedule("DHGR", n)
    fractions = (1.0, 0.5, 0.1, 0.02)
    clips = np.stack([synth.synthetic_frames("DHGR", n, fractions[c % 4], seed=5000 + c)
                      for c in range(n_clips)])
    seeds = [900 + c for c in range(n_clips)]
    got, info, _ = run_device(ops, "DHGR", device_tables("DHGR"), clips, segs, seeds)
    for c in range(n_clips):
        want, real, *_ = run_oracle("DHGR", tab, clips[c], segs, seeds[c])
        assert np.array_equal(got[c][:, :6].astype(np.int64), want), "clip %d" % c
        assert np.array_equal(got[c][:, 6], real), "clip %d" % c


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_score_frames_equals_pack_diff_weights_and_priority_fold(mode, device_tables):
    """iiv_score_frames = _pack + diff_weights per bank + video.py:111, :115-116."""
    import torch
    from iivision_b200 import ops, synth
    table = device_tables(mode)
    n = 6
    fr = synth.synthetic_frames(mode, n + 1, 0.3, seed=77)
    banks = fr.shape[1]
    d = torch.from_numpy(fr).cuda()
    packed = ops.pack(mode, d[:, 0].contiguous(), d[:, 1].contiguous() if banks == 2 else None)
    src, tgt_mem = packed[:n].contiguous(), d[1:].contiguous()
    rng = np.random.default_rng(3)
    prio0 = rng.integers(0, 5000, size=(n, banks, 32, 256)).astype(np.int32)
    prio0[rng.random(prio0.shape) < 0.3] = 0
    prio = torch.from_numpy(prio0.copy()).cuda()
    tp, diff = ops.score_frames(mode, src, tgt_mem, table, priority=prio)
    assert torch.equal(tp, packed[1:])
    holes = synth.screen_holes()
    for b in range(banks):
        want = ops.diff_weights(mode, b, src, packed[1:].contiguous(), table).cpu().numpy()
        want[:, holes] = 0
        assert np.array_equal(diff[:, b].cpu().numpy(), want)
        p = prio0[:, b].copy()
        p[want == 0] = 0
        p += want
        assert np.array_equal(prio[:, b].cpu().numpy(), p)
    # one source bitmap for every frame, holes kept, no priorities
    tp2, diff2 = ops.score_frames(mode, src[0], tgt_mem, table, zero_holes=False)
    for b in range(banks):
        want = ops.diff_weights(mode, b, src[0:1].expand(n, 32, 128).contiguous(),
                                packed[1:].contiguous(), table)
        assert torch.equal(diff2[:, b], want)

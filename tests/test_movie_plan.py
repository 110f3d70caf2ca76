"""CPU: the host-side schedule of a whole movie (movie.plan_movie) against what the
unmodified reference's Movie.encode + Movie.emit_stream did on the same inputs
(tests/golden/movie_*.npz, generator oracle/make_golden.py)."""

import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "movie_*.npz")))


@pytest.fixture(scope="module")
def movie():
    from iivision_b200 import _build
    _build.build()
    from iivision_b200 import movie
    return movie


def _plan(movie, g):
    return movie.plan_movie(str(g["mode"]), len(g["audio"]), g["frames"].shape[0],
                            float(g["sample_rate"]), float(g["input_frame_rate"]),
                            int(g["every_n_video_frames"]), int(g["max_bytes_out"]) or None)


def test_cases_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("case", CASES)
def test_plan_matches_reference_run(movie, case):
    from iivision_b200._lib import lib
    g = np.load(os.path.join(GOLDEN, "movie_%s.npz" % case))
    plan = _plan(movie, g)
    data = g["bytes"]
    # stream length: header + emitted tick opcodes + Acks + Terminate + padding
    assert int(lib.iiv_stream_length(plan.emitted, 1)) == len(data)
    assert sum(s[2] for s in plan.segments) == plan.pulled
    assert plan.emitted in (plan.pulled, plan.pulled - 1)
    # Movie.ticks counts the tick whose frame could not be grabbed as well
    assert int(g["ticks_pulled"]) in (plan.pulled, plan.pulled + 1)
    # the speaker ticks sit in the opcode addresses; check the bank the Acks select instead:
    # every 2 KiB frame ends in (ack address, $54 | aux, $ff), the bank alternating in DHGR
    acks = [int(data[k * 2048 + 2046]) for k in range(len(data) // 2048 - 1)]
    if str(g["mode"]) == "DHGR":
        assert acks == [0x55 if k % 2 == 0 else 0x54 for k in range(len(acks))]
    else:
        assert set(acks) <= {0x54}
    # generators: new frame or bank flip, never empty, banks as the Acks left them
    for (slot, is_aux, count) in plan.segments:
        assert count > 0 and 0 <= slot < len(plan.frames_used)
        assert is_aux in (0, 1) and (str(g["mode"]) == "DHGR" or is_aux == 0)
    every = int(g["every_n_video_frames"])
    assert plan.frames_used == list(range(0, g["frames"].shape[0], every))[:len(plan.frames_used)]


def test_plan_degenerate_inputs(movie):
    assert movie.plan_movie("DHGR", 0, 3).pulled == 0
    assert movie.plan_movie("HGR", 100, 0) == movie.MoviePlan([], [], 0, 0)
    p = movie.plan_movie("DHGR", 1000, 1, 14700., 30., 1, None)
    # a single frame: its successor is due at tick 490 and cannot be grabbed
    assert p.pulled == 489 and p.frames_used == [0]
    assert p.segments == [(0, 0, 291), (0, 1, 198)]
    # max_bytes_out below the header: the first opcode is computed and dropped
    q = movie.plan_movie("HGR", 50, 1, max_bytes_out=5)
    assert (q.pulled, q.emitted) == (1, 0)


def test_movie_needs_media_sources(movie):
    with pytest.raises(NotImplementedError):
        movie.Movie("clip.mp4")


def test_pull_predictor_learns_the_movie_schedule(movie):
    """video.PullPredictor fed the generators of plan_movie: once one bank-flip interval and
    one frame have been seen, every guess is exactly the number of pulls that follows."""
    from iivision_b200 import video
    plan = movie.plan_movie("DHGR", 490 * 12, 12, 14700., 30., 2)
    pred = video.PullPredictor(292, True, video.MAX_BUDGET)
    targets = {slot: object() for slot in range(len(plan.frames_used))}
    misses = []
    for k, (slot, is_aux, count) in enumerate(plan.segments):
        g = pred.guess(targets[slot], bool(is_aux))
        assert 1 <= g <= video.MAX_BUDGET
        if g != count:
            misses.append(k)
        pred.pulled(count)
    first_of_third_frame = next(k for k, s_ in enumerate(plan.segments) if s_[0] == 2)
    # the first 2 KiB stream frame is one opcode short (header) and so is the first video
    # frame (tick numbering starts at 1): all learnt by the time the third frame starts
    assert misses and max(misses) < first_of_third_frame, misses
    # HGR: no flips, one generator per frame
    plan = movie.plan_movie("HGR", 490 * 8, 8, 14700., 30., 2)
    pred = video.PullPredictor(980, False, video.MAX_BUDGET)
    guesses = []
    for slot, is_aux, count in plan.segments:
        guesses.append((pred.guess(object(), False), count))
        pred.pulled(count)
    assert all(g == c for g, c in guesses[2:-1]), guesses


@pytest.mark.parametrize("case", CASES)
def test_plan_plus_oracle_encoder_reproduces_reference_bytes(movie, oracle_tables, case):
    """The whole movie on the CPU: plan_movie's generators run through the numpy oracle of
    the encoder, the tuples through the host muxer -- against the bytes the unmodified
    reference's Movie.encode + emit_stream produced.  Pins the schedule (which opcode comes
    from which frame/bank generator) and the oracle's encoder on a full-length run."""
    import random
    from iivision_b200 import opcodes, video_mode
    from oracle import scorer
    b = np.load(os.path.join(GOLDEN, "byte_stream.npz"))
    opcodes.set_addresses(b["tick_addr"], int(b["ack_addr"]), int(b["terminate_addr"]))
    g = np.load(os.path.join(GOLDEN, "movie_%s.npz" % case))
    mode = str(g["mode"])
    frames, samples = g["frames"], g["audio"]
    plan = _plan(movie, g)
    seed = int(g["rng_seed"])
    v = scorer.OracleVideo(mode, oracle_tables(mode), py_rng=random.Random(seed),
                           np_rng=np.random.RandomState(seed))
    vm = getattr(video_mode.VideoMode, mode)
    ops_ = [opcodes.Header(mode=vm)]
    tick = 0
    for slot, is_aux, count in plan.segments:
        fr = plan.frames_used[slot]
        tgt = v.target_bitmap(frames[fr, 0], frames[fr, 1] if mode == "DHGR" else None)
        seq = v.encode_frame(tgt, bool(is_aux))
        for _ in range(count):
            page, content, offs = next(seq)
            if tick < plan.emitted:
                ops_.append(opcodes.TICK_OPCODES[(2 * int(samples[tick]) + 34, int(page))](
                    int(content), tuple(int(o) for o in offs)))
            tick += 1
    assert tick == plan.pulled
    mux = movie.StreamMuxer(vm, max_bytes_out=int(g["max_bytes_out"]) or None)
    data = bytes(mux.emit_stream(ops_))
    assert data == g["bytes"].tobytes()
    assert np.array_equal(v.pixelmap.packed, g["packed"])
    assert np.array_equal(v.update_priority, g["priority_main"])


def test_plan_properties_random(movie):
    """plan_movie against the C library's stream arithmetic and its own invariants on random
    parameters (hypothesis is not needed for this: a seeded sweep)."""
    from iivision_b200._lib import lib
    rng = np.random.default_rng(17)
    for _ in range(300):
        mode = "DHGR" if rng.random() < 0.6 else "HGR"
        n_samples = int(rng.integers(0, 6000))
        n_frames = int(rng.integers(0, 12))
        fps = float(rng.choice([23.976, 24., 25., 29.97, 30., 60.]))
        every = int(rng.integers(1, 4))
        cap = int(rng.integers(1, 40000)) if rng.random() < 0.5 else None
        free = movie.plan_movie(mode, n_samples, n_frames, 14700., fps, every, None)
        plan = movie.plan_movie(mode, n_samples, n_frames, 14700., fps, every, cap)
        assert free.emitted == free.pulled <= n_samples
        assert sum(s[2] for s in plan.segments) == plan.pulled
        assert all(s[2] > 0 for s in plan.segments)
        if cap is None:
            assert plan == free
            continue
        within = int(lib.iiv_stream_ticks_within(free.emitted, cap))
        assert plan.emitted == within
        # the opcode that finds the stream full is computed, then dropped
        assert plan.pulled == within + (1 if within < free.pulled else 0)
        # a capped plan is a prefix of the free one
        flat_free = [(s[0], s[1]) for s in free.segments for _ in range(s[2])]
        flat_cap = [(s[0], s[1]) for s in plan.segments for _ in range(s[2])]
        assert flat_cap == flat_free[:len(flat_cap)]
        if mode == "HGR":
            assert all(s[1] == 0 for s in plan.segments)
        else:
            # the bank of tick k is decided by how many Acks precede it: 291, then every 292
            for k, (_, is_aux) in enumerate(flat_free[:2000:37]):
                kk = k * 37
                flips = 0 if kk < 291 else 1 + (kk - 291) // 292
                assert is_aux == flips % 2

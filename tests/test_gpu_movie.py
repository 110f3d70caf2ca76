"""GPU: the whole Movie.encode + Movie.emit_stream path (SURVEY "next" row N2) against
byte streams produced by the unmodified reference on the same synthetic frames, audio
samples and seeds (tests/golden/movie_*.npz, generator oracle/make_golden.py) -- once
through the reference-named generator API (movie.Movie) and once as two kernel launches
(movie.transcode_device)."""

import contextlib
import glob
import io
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "movie_*.npz")))


@pytest.fixture(scope="module")
def mods():
    import types
    from iivision_b200 import movie, opcodes, ops, palette, screen, video_mode
    b = np.load(os.path.join(GOLDEN, "byte_stream.npz"))
    opcodes.set_addresses(b["tick_addr"], int(b["ack_addr"]), int(b["terminate_addr"]))
    return types.SimpleNamespace(movie=movie, opcodes=opcodes, ops=ops, palette=palette,
                                 screen=screen, video_mode=video_mode)


def _check_state(g, mode, packed, main, prio_main, aux=None, prio_aux=None):
    assert np.array_equal(packed, g["packed"])
    assert np.array_equal(main, g["main"])
    assert np.array_equal(prio_main, g["priority_main"])
    if mode == "DHGR":
        assert np.array_equal(aux, g["aux"])
        assert np.array_equal(prio_aux, g["priority_aux"])


@pytest.mark.parametrize("case", CASES)
def test_movie_generator_api(mods, case):
    g = np.load(os.path.join(GOLDEN, "movie_%s.npz" % case))
    mode = str(g["mode"])
    frames, samples = g["frames"], g["audio"]

    class Audio:
        sample_rate = float(g["sample_rate"])

        @staticmethod
        def audio_stream():
            yield from (int(a) for a in samples)

    class Grabber:
        input_frame_rate = float(g["input_frame_rate"])

        @staticmethod
        def frames():
            for k in range(frames.shape[0]):
                yield (mods.screen.MemoryMap(1, frames[k, 0].copy()),
                       mods.screen.MemoryMap(1, frames[k, 1].copy()) if mode == "DHGR" else None)

    random.seed(int(g["rng_seed"]))
    np.random.seed(int(g["rng_seed"]))
    m = mods.movie.Movie("synthetic", every_n_video_frames=int(g["every_n_video_frames"]),
                         max_bytes_out=int(g["max_bytes_out"]) or None,
                         video_mode=getattr(mods.video_mode.VideoMode, mode),
                         palette=mods.palette.Palette.NTSC, audio=Audio(), frame_grabber=Grabber())
    with contextlib.redirect_stdout(io.StringIO()):
        data = bytes(m.emit_stream(m.encode()))
    assert np.array_equal(np.frombuffer(data, np.uint8), g["bytes"])
    assert m.ticks == int(g["ticks_pulled"])
    m.video.sync()
    v = m.video
    _check_state(g, mode, v.pixelmap.packed, v.memory_map.page_offset, v.update_priority,
                 getattr(v, "aux_memory_map", v.memory_map).page_offset,
                 getattr(v, "aux_update_priority", None))


@pytest.mark.parametrize("case", CASES)
def test_transcode_device(mods, device_tables, case):
    import torch
    from encoder_util import seed_states
    ops = mods.ops
    g = np.load(os.path.join(GOLDEN, "movie_%s.npz" % case))
    mode = str(g["mode"])
    b = np.load(os.path.join(GOLDEN, "byte_stream.npz"))
    states = ops.new_clip_states(1)
    seed_states(ops, states, [int(g["rng_seed"])])
    data, plan = mods.movie.transcode_device(
        mode, g["frames"], g["audio"], device_tables(mode), states,
        sample_rate=float(g["sample_rate"]), input_frame_rate=float(g["input_frame_rate"]),
        every_n_video_frames=int(g["every_n_video_frames"]),
        max_bytes_out=int(g["max_bytes_out"]) or None,
        addresses=(b["tick_addr"], int(b["ack_addr"]), int(b["terminate_addr"])))
    assert np.array_equal(data.cpu().numpy(), g["bytes"])

    def field(f, dtype, shape):
        return ops.state_field(states, f, dtype, shape)[0].cpu().numpy()
    _check_state(g, mode, field(ops.F_PACKED, torch.int64, (32, 128)).view(np.uint64),
                 field(ops.F_MAIN, torch.uint8, (32, 256)),
                 field(ops.F_PRIO_MAIN, torch.int32, (32, 256)),
                 field(ops.F_AUX, torch.uint8, (32, 256)),
                 field(ops.F_PRIO_AUX, torch.int32, (32, 256)))


def test_movie_from_the_frame_cache(mods, tmp_path):
    """The same movie with its frames served by frame_grabber.FileFrameGrabber from the
    reference's on-disk cache of converted frames (frame_grabber.py:73-76, :95-99): the bytes
    of the reference's Movie.encode + emit_stream."""
    from iivision_b200 import frame_grabber
    case = [c for c in CASES if "DHGR" in c.upper()] or CASES
    g = np.load(os.path.join(GOLDEN, "movie_%s.npz" % case[0]))
    mode = str(g["mode"])
    frames, samples = g["frames"], g["audio"]
    vm = getattr(mods.video_mode.VideoMode, mode)
    video_file = str(tmp_path / "synthetic.mp4")
    d = frame_grabber.FileFrameGrabber._output_dir(video_file, vm, mods.palette.Palette.NTSC)
    os.makedirs(d)
    for k in range(frames.shape[0]):
        if mode == "DHGR":
            frames[k, 0].reshape(8192).tofile("%s/%08d.BIN" % (d, k))
            frames[k, 1].reshape(8192).tofile("%s/%08d.AUX" % (d, k))
        else:
            frames[k, 0].reshape(8192).tofile("%s/%08dC.BIN" % (d, k))

    class Audio:
        sample_rate = float(g["sample_rate"])

        @staticmethod
        def audio_stream():
            yield from (int(a) for a in samples)

    grabber = frame_grabber.FileFrameGrabber(video_file, vm, mods.palette.Palette.NTSC,
                                             input_frame_rate=float(g["input_frame_rate"]))
    assert np.array_equal(grabber.frames_array(), frames)
    random.seed(int(g["rng_seed"]))
    np.random.seed(int(g["rng_seed"]))
    m = mods.movie.Movie(video_file, every_n_video_frames=int(g["every_n_video_frames"]),
                         max_bytes_out=int(g["max_bytes_out"]) or None, video_mode=vm,
                         palette=mods.palette.Palette.NTSC, audio=Audio(), frame_grabber=grabber)
    with contextlib.redirect_stdout(io.StringIO()):
        data = bytes(m.emit_stream(m.encode()))
    assert np.array_equal(np.frombuffer(data, np.uint8), g["bytes"])

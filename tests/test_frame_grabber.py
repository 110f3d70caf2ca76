"""iivision_b200.frame_grabber: the reference's directory convention and its cache of
converted frames (frame_grabber.py:56-147), read back as the reference reads it."""

import os
import sys

import numpy as np
import pytest

from iivision_b200 import frame_grabber, screen
from iivision_b200.palette import Palette
from iivision_b200.video_mode import VideoMode

REF = "/root/reference/transcoder"


def _write_cache(tmp_path, mode, n, seed=0):
    rng = np.random.default_rng(seed)
    video = str(tmp_path / "clip.v1.mp4")
    d = frame_grabber.FileFrameGrabber._output_dir(video, mode, Palette.NTSC)
    os.makedirs(d)
    frames = rng.integers(0, 256, size=(n, 2 if mode == VideoMode.DHGR else 1, 8192),
                          dtype=np.uint8)
    for k in range(n):
        if mode == VideoMode.DHGR:
            frames[k, 0].tofile("%s/%08d.BIN" % (d, k))
            frames[k, 1].tofile("%s/%08d.AUX" % (d, k))
        else:
            frames[k, 0].tofile("%s/%08dC.BIN" % (d, k))
    return video, d, frames


def test_output_dir_convention():
    # the literals of the reference's frame_grabber_test.py
    f = frame_grabber.FileFrameGrabber._output_dir
    assert f("/foo/bar.mp4", VideoMode.DHGR, Palette.NTSC) == "/foo/bar/DHGR/NTSC"
    assert f("/foo/bar.blee.mp4", VideoMode.HGR, Palette.IIGS) == "/foo/bar.blee/HGR/IIGS"
    assert f("/foo/bar blee.mp4", VideoMode.DHGR, Palette.IIGS) == "/foo/bar blee/DHGR/IIGS"


@pytest.mark.parametrize("mode", [VideoMode.HGR, VideoMode.DHGR])
def test_cached_frames_round_trip(tmp_path, mode):
    video, d, frames = _write_cache(tmp_path, mode, 5)
    g = frame_grabber.FileFrameGrabber(video, mode, Palette.NTSC, input_frame_rate=24.0)
    assert g.video_mode == mode and g.input_frame_rate == 24.0
    got = list(g.frames())
    assert len(got) == 5
    for k, (main, aux) in enumerate(got):
        assert isinstance(main, screen.MemoryMap) and main.screen_page == 1
        # FlatMemoryMap(...).to_memory_map(): address 0x2000 + 256 * page + offset
        assert np.array_equal(main.page_offset, frames[k, 0].reshape(32, 256))
        if mode == VideoMode.DHGR:
            assert np.array_equal(aux.page_offset, frames[k, 1].reshape(32, 256))
        else:
            assert aux is None
    arr = g.frames_array()
    assert arr.shape == (5, frames.shape[1], 32, 256)
    assert np.array_equal(arr.reshape(frames.shape), frames)


def test_missing_and_broken_caches(tmp_path):
    g = frame_grabber.FileFrameGrabber(str(tmp_path / "none.mp4"), VideoMode.DHGR, Palette.NTSC)
    with pytest.raises(FileNotFoundError, match="bmp2dhr"):
        next(g.frames())
    video, d, _ = _write_cache(tmp_path, VideoMode.DHGR, 3)
    os.remove("%s/%08d.AUX" % (d, 2))
    with pytest.raises(FileNotFoundError, match="other bank"):
        list(frame_grabber.FileFrameGrabber(video, VideoMode.DHGR, Palette.NTSC).frames())
    with open("%s/%08d.AUX" % (d, 2), "wb") as f:
        f.write(b"\0" * 100)
    with pytest.raises(ValueError, match="8192"):
        list(frame_grabber.FileFrameGrabber(video, VideoMode.DHGR, Palette.NTSC).frames())
    with pytest.raises(NotImplementedError):
        frame_grabber.FrameGrabber(VideoMode.HGR).frames()


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
@pytest.mark.parametrize("mode_name", ["HGR", "DHGR"])
def test_same_memory_maps_as_the_reference_reads(tmp_path, mode_name):
    """The unmodified reference's cache-hit path (frame_grabber.py:73-76, :95-99, :134-145:
    np.fromfile + FlatMemoryMap.to_memory_map) on the same files."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import ref_harness
    ref = ref_harness.load()
    mode = VideoMode[mode_name]
    video, d, _ = _write_cache(tmp_path, mode, 4, seed=3)
    ours = list(frame_grabber.FileFrameGrabber(video, mode, Palette.NTSC).frames())
    ref_dir = ref.frame_grabber.FileFrameGrabber._output_dir(
        video, ref.video_mode.VideoMode[mode_name], ref.palette.Palette.NTSC)
    assert ref_dir == d
    for k, (main, aux) in enumerate(ours):
        names = (["%s/%08d.BIN" % (d, k), "%s/%08d.AUX" % (d, k)] if mode_name == "DHGR"
                 else ["%s/%08dC.BIN" % (d, k)])
        want = [ref.screen.FlatMemoryMap(screen_page=1, data=np.fromfile(n, dtype=np.uint8))
                .to_memory_map() for n in names]
        assert np.array_equal(main.page_offset, want[0].page_offset)
        if mode_name == "DHGR":
            assert np.array_equal(aux.page_offset, want[1].page_offset)

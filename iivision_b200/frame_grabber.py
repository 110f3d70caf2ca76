"""Extracts sequence of still images from input video stream (reference
transcoder/frame_grabber.py) -- the half of it that needs no external tool.

The reference's ``FileFrameGrabber.frames()`` (frame_grabber.py:56-147) decodes the video
with ffmpeg, resizes each frame with PIL and converts it with the ``bmp2dhr`` binary, and it
keeps every converted frame on disk: ``<video without extension>/<MODE>/<PALETTE>/
%08dC.BIN`` (HGR) or ``%08d.BIN`` + ``%08d.AUX`` (DHGR), 8 KiB of screen memory each, which
a later run reads back instead of converting again (:73-76, :95-99).  None of ffmpeg, PIL's
encoder path or bmp2dhr exists in this package's environment (row N3),
so what is provided is

* ``FrameGrabber``: the interface ``video.Video`` / ``movie.Movie`` consume
  (``video_mode``, ``input_frame_rate``, ``frames()``), as in the reference (:18-24);
* ``FileFrameGrabber``: the reference's class name and directory convention; its ``frames()``
  serves the converted frames a previous run of the reference (or bmp2dhr by hand) left in
  that directory, in the reference's order and as the reference's ``(main, aux)``
  ``MemoryMap`` pairs, and says which tool is missing when a frame is not there;
* ``frames_array()``: the same frames as one ``uint8[n, banks, 32, 256]`` array, the input of
  ``movie.transcode_device`` / ``ops.encode_clips``.
"""

import os
from typing import Iterator, Optional, Tuple

import numpy as np

from . import screen
from .palette import Palette
from .video_mode import VideoMode


class FrameGrabber:
    def __init__(self, mode: VideoMode):
        self.video_mode = mode
        self.input_frame_rate = 30

    def frames(self) -> Iterator[screen.MemoryMap]:
        raise NotImplementedError


class FileFrameGrabber(FrameGrabber):
    """Frames of ``filename`` as (D)HGR screen memory, from the reference's on-disk cache of
    converted frames.  ``input_frame_rate`` cannot be probed without ffmpeg: pass the
    source's rate (the reference reads ``r_frame_rate``, frame_grabber.py:35-39)."""

    def __init__(self, filename, mode: VideoMode, palette: Palette,
                 input_frame_rate: float = 30.0):
        super(FileFrameGrabber, self).__init__(mode)
        self.filename = filename  # type: str
        self.palette = palette  # type: Palette
        self.input_frame_rate = float(input_frame_rate)

    @staticmethod
    def _output_dir(filename, video_mode, palette) -> str:
        return "%s/%s/%s" % (
            ".".join(filename.split(".")[:-1]),
            video_mode.name,
            palette.name)

    def _frame_files(self, idx: int) -> Tuple[str, Optional[str]]:
        frame_dir = self._output_dir(self.filename, self.video_mode, self.palette)
        if self.video_mode == VideoMode.DHGR:
            return ("%s/%08d.BIN" % (frame_dir, idx), "%s/%08d.AUX" % (frame_dir, idx))
        return ("%s/%08dC.BIN" % (frame_dir, idx), None)

    @staticmethod
    def _read_bank(path: str) -> np.ndarray:
        data = np.fromfile(path, dtype=np.uint8)
        if data.shape != (8192,):
            raise ValueError("%s: expected 8192 bytes of screen memory, found %d"
                             % (path, data.size))
        return data

    def _banks(self) -> Iterator[Tuple[np.ndarray, Optional[np.ndarray]]]:
        idx = 0
        while True:
            main_file, aux_file = self._frame_files(idx)
            have_main = os.path.exists(main_file)
            have_aux = aux_file is None or os.path.exists(aux_file)
            if not have_main and not (aux_file and os.path.exists(aux_file)):
                if idx == 0:
                    raise FileNotFoundError(
                        "%s: no converted frames.  Converting %r needs ffmpeg (scikit-video), "
                        "PIL and /usr/local/bin/bmp2dhr (frame_grabber.py:56-113), which this "
                        "package does not provide; run the reference's frame grabber once, or "
                        "bmp2dhr by hand, to fill that directory."
                        % (os.path.dirname(main_file), self.filename))
                return
            if not (have_main and have_aux):
                raise FileNotFoundError("frame %d: %s without its other bank"
                                        % (idx, main_file if have_main else aux_file))
            yield (self._read_bank(main_file),
                   self._read_bank(aux_file) if aux_file else None)
            idx += 1

    def frames(self) -> Iterator[screen.MemoryMap]:
        """(main, aux) MemoryMaps per frame (aux None for HGR), frame_grabber.py:134-145."""
        for main, aux in self._banks():
            main_map = screen.FlatMemoryMap(screen_page=1, data=main).to_memory_map()
            if aux is None:
                aux_map = None
            else:
                aux_map = screen.FlatMemoryMap(screen_page=1, data=aux).to_memory_map()
            yield (main_map, aux_map)

    def frames_array(self) -> np.ndarray:
        """Every frame as uint8[n, banks, 32, 256] (banks: main[, aux]): what
        movie.transcode_device and ops.encode_clips take."""
        out = [np.stack([b.reshape(32, 256) for b in banks if b is not None])
               for banks in self._banks()]
        return np.stack(out)

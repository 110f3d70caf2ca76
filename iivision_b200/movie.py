"""Multiplexing of video and audio into the player's byte stream: the reference's
transcoder/movie.py (SURVEY.md "next" row N2) on top of the CUDA encoder.

* ``Movie``: the reference class with its tick loop (``encode``, movie.py:56-111) and
  its generator-based byte emitter (``emit_stream`` / ``done``, movie.py:122-161).  The
  two media front ends -- audio decoding (audioread + librosa) and the frame grabber
  (ffmpeg + bmp2dhr) -- are not part of this package: ``Movie`` takes any objects with
  the reference's interfaces (``audio.sample_rate`` / ``audio.audio_stream()``,
  ``frame_grabber.input_frame_rate`` / ``frame_grabber.frames()``) and fails loudly
  without them.
* ``StreamMuxer``: ``emit_stream`` / ``done`` alone, for opcode streams made elsewhere.
* ``plan_movie`` + ``transcode_device``: the same movie in two kernel launches.  Which
  opcode belongs to which (frame, bank) generator is a function of the tick count, the
  frame rate and the stream position alone -- never of the opcodes' contents -- so the
  whole schedule is worked out on the host first, ``iiv_encode_clips`` runs it, and
  ``iiv_emit_stream`` lays out header, tick opcodes, Acks, Terminate and padding.
* ``emit_stream_device(...)``: the byte layout kernel on its own.
* ``stream_schedule(...)``: the fixed-rate special case of ``plan_movie``.
"""

from typing import Iterable, Iterator, List, NamedTuple, Optional, Tuple

import numpy as np
import torch

from . import ops
from . import opcodes
from . import screen
from . import video
from ._lib import check, lib
from .machine import Machine
from .palette import Palette
from .video_mode import VideoMode


class StreamMuxer:
    """emit_stream / done of the reference's Movie, without the media front end."""

    def __init__(self, video_mode: VideoMode = VideoMode.HGR, max_bytes_out: int = None):
        self.video_mode = video_mode
        self.max_bytes_out = max_bytes_out
        self.stream_pos = 0
        self.state = Machine()
        self.aux_memory_bank = False

    def _emit_bytes(self, _op: opcodes.Opcode) -> Iterable[int]:
        for b in self.state.emit(_op):
            yield b
            self.stream_pos += 1

    def emit_stream(self, ops_: Iterable[opcodes.Opcode]) -> Iterator[int]:
        """Compiled byte stream of an opcode stream, with the Ack opcodes that close
        every 2 KiB TCP frame (and flip MAIN/AUX in DHGR)."""
        for op in ops_:
            if self.max_bytes_out and self.stream_pos >= self.max_bytes_out:
                yield from self.done()
                return
            yield from self._emit_bytes(op)
            if self.stream_pos % 2048 >= 2044:
                if self.video_mode == VideoMode.DHGR:
                    self.aux_memory_bank = not self.aux_memory_bank
                yield from self._emit_bytes(opcodes.Ack(self.aux_memory_bank))
                assert self.stream_pos % 2048 == 0, self.stream_pos % 2048
        yield from self.done()

    def done(self) -> Iterator[int]:
        """Terminal opcode, then zero padding to the 2 KiB boundary."""
        yield from self._emit_bytes(opcodes.Terminate())
        for _ in range(2048 - (self.stream_pos % 2048)):
            yield 0x00


class Movie(StreamMuxer):
    """The reference's ``movie.Movie`` with the media sources passed in.

    ``audio`` must offer ``sample_rate`` and ``audio_stream()`` (speaker samples -15..16,
    audio.py:84-103); ``frame_grabber`` must offer ``input_frame_rate`` and ``frames()``
    yielding ``(main, aux)`` ``screen.MemoryMap`` pairs (``aux`` is None for HGR;
    frame_grabber.py:56-140).
    """

    def __init__(self, filename: str = None, every_n_video_frames: int = 1,
                 audio_bitrate: int = 14700, audio_normalization: float = None,
                 max_bytes_out: int = None, video_mode: VideoMode = VideoMode.HGR,
                 palette: Palette = Palette.NTSC, *, audio=None, frame_grabber=None):
        super().__init__(video_mode, max_bytes_out)
        if audio is None or frame_grabber is None:
            raise NotImplementedError(
                "decoding %r needs the reference's audio.Audio and "
                "frame_grabber.FileFrameGrabber (audioread, librosa, ffmpeg, bmp2dhr), which "
                "this package does not provide: pass audio= and frame_grabber=" % (filename,))
        self.filename = filename
        self.every_n_video_frames = every_n_video_frames
        self.palette = palette
        self.audio = audio
        self.frame_grabber = frame_grabber
        self.video = video.Video(frame_grabber, ticks_per_second=audio.sample_rate,
                                 mode=video_mode, palette=palette)
        self.ticks = 0      # audio tick opcodes so far

    def _target(self, main, aux):
        if self.video_mode == VideoMode.DHGR:
            return screen.DHGRBitmap(main_memory=main, aux_memory=aux, palette=self.palette)
        return screen.HGRBitmap(main_memory=main, palette=self.palette)

    def encode(self) -> Iterator[opcodes.Opcode]:
        """One tick opcode per audio sample; the video generator behind it is replaced at
        every encoded frame and whenever ``emit_stream`` has flipped the memory bank."""
        grabbed = self.frame_grabber.frames()
        yield opcodes.Header(mode=self.video_mode)
        target, op_seq = None, None
        bank_in_use = self.aux_memory_bank
        for sample in self.audio.audio_stream():
            self.ticks += 1
            if self.video.tick(self.ticks):
                try:
                    main, aux = next(grabbed)
                except StopIteration:
                    break
                if (self.video.frame_number - 1) % self.every_n_video_frames == 0:
                    target = self._target(main, aux)
                    print("Starting frame %d" % self.video.frame_number)
                    op_seq = self.video.encode_frame(target, is_aux=self.aux_memory_bank)
                    self.video.out_of_work = {True: False, False: False}
            if bank_in_use != self.aux_memory_bank:
                bank_in_use = self.aux_memory_bank
                op_seq = self.video.encode_frame(target, is_aux=bank_in_use)
            page, content, offsets = next(op_seq)
            # samples -15..16 -> speaker duty cycles 4..66 in steps of 2
            yield opcodes.TICK_OPCODES[(2 * int(sample) + 34, page)](content, offsets)


class MoviePlan(NamedTuple):
    segments: List[Tuple[int, int, int]]   # (index into frames_used, is_aux, opcodes pulled)
    frames_used: List[int]                 # grabber frame numbers that get encoded
    pulled: int                            # tick opcodes computed (encode_frame pulls)
    emitted: int                           # tick opcodes that reach the byte stream


def plan_movie(mode: str, n_samples: int, n_frames: int, sample_rate: float = 14700.,
               input_frame_rate: float = 30., every_n_video_frames: int = 1,
               max_bytes_out: Optional[int] = None) -> MoviePlan:
    """Who pulls what in ``Movie.emit_stream(Movie.encode())``, without running it.

    Follows the two interleaved generators tick by tick: ``Video.tick`` decides when a
    frame is due (ticks >= ticks_per_frame * frame_number in float arithmetic,
    video.py:64-70), every ``every_n_video_frames``-th grabbed frame starts a generator,
    the byte position after each 7-byte tick opcode decides when an Ack closes the 2 KiB
    frame and, in DHGR, flips the bank for the next tick (movie.py:139-150).  The stream
    ends when the samples or the frames run out, or at the first opcode pulled with the
    position at or past ``max_bytes_out`` (that opcode is computed but never emitted,
    movie.py:132-134).
    """
    dhgr = ops.mode_id(mode) == ops.MODE_DHGR
    ticks_per_frame = float(sample_rate) / input_frame_rate
    frame_number = 0
    grabbed = 0
    aux, bank_in_use = False, False
    pos = 7                     # the header is as long as a tick opcode
    segments, frames_used = [], []
    cur = None                  # [frames_used index, is_aux, count] of the live generator
    pulled = emitted = 0

    def start(frame_slot):
        nonlocal cur
        if cur is not None and cur[2] > 0:
            segments.append(tuple(cur))
        cur = [frame_slot, int(aux), 0]

    for tick in range(1, n_samples + 1):
        if tick >= ticks_per_frame * frame_number:
            frame_number += 1
            if grabbed == n_frames:
                break
            grabbed += 1
            if (frame_number - 1) % every_n_video_frames == 0:
                frames_used.append(grabbed - 1)
                start(len(frames_used) - 1)
        if bank_in_use != aux:
            bank_in_use = aux
            start(cur[0])
        cur[2] += 1
        pulled += 1
        if max_bytes_out and pos >= max_bytes_out:
            break
        emitted += 1
        pos += 7
        if pos % 2048 >= 2044:
            if dhgr:
                aux = not aux
            pos += 4
    if cur is not None and cur[2] > 0:
        segments.append(tuple(cur))
    return MoviePlan(segments, frames_used, pulled, emitted)


def transcode_device(mode: str, frames: np.ndarray, samples: np.ndarray, table: torch.Tensor,
                     states: torch.Tensor, sample_rate: float = 14700.,
                     input_frame_rate: float = 30., every_n_video_frames: int = 1,
                     max_bytes_out: Optional[int] = None, addresses=None
                     ) -> Tuple[torch.Tensor, MoviePlan]:
    """A whole movie in one encoder launch and one byte-layout launch.

    frames: uint8[n_frames, banks, 32, 256] memory maps as the frame grabber would yield
    them; samples: speaker samples -15..16, one per tick; table: the symmetric edit-distance
    table of the mode/palette; states: one clip state blob (``ops.new_clip_states(1)`` with
    both MT19937 fields set), updated in place to where the reference's ``Video`` and global
    generators stand when ``Movie.encode`` stops.  Returns the ``.a2m`` bytes (uint8 tensor
    on the device) and the plan.
    """
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    samples = np.asarray(samples)
    plan = plan_movie(mode, len(samples), frames.shape[0], sample_rate, input_frame_rate,
                      every_n_video_frames, max_bytes_out)
    m = ops.mode_id(mode)
    used = frames[plan.frames_used] if plan.frames_used else frames[:0]
    if plan.pulled:
        tmem = torch.from_numpy(used[None]).cuda()
        flat = tmem.view(-1, used.shape[1], 32, 256)
        packed = ops.pack(mode, flat[:, 0].contiguous(),
                          flat[:, 1].contiguous() if m == ops.MODE_DHGR else None)
        records, _ = ops.encode_clips(mode, states, tmem,
                                      packed.view(1, len(plan.frames_used), 32, 128),
                                      plan.segments, table)
        records = records[0]
    else:
        records = torch.zeros((0, 8), dtype=torch.uint8, device="cuda")
    ticks = torch.from_numpy(
        (2 * samples[:plan.emitted].astype(np.int64) + 34).astype(np.uint8)).cuda()
    data = emit_stream_device(mode, records[:plan.emitted].contiguous(), ticks, None,
                              addresses)
    return data, plan


def emit_stream_device(mode, opcode_records: torch.Tensor, ticks: torch.Tensor,
                       max_bytes_out: int = None, addresses=None) -> torch.Tensor:
    """Header + tick opcodes + Acks + Terminate + padding as one uint8 tensor.

    opcode_records: uint8[n, 8] as written by ``ops.encode_clips`` (page + 32, content,
    four offsets, ...); ticks: uint8[n] speaker ticks (4..66, even).  ``addresses`` =
    (uint16[32][32], ack, terminate), default ``opcodes.address_table()``.
    """
    m = ops.mode_id(mode)
    table, ack, terminate = addresses if addresses is not None else opcodes.address_table()
    n = int(opcode_records.shape[0])
    if opcode_records.shape != (n, 8) or ticks.shape != (n,):
        raise ValueError("opcode_records must be uint8[n, 8] and ticks uint8[n]")
    n_emit = int(lib.iiv_stream_ticks_within(n, int(max_bytes_out or 0)))
    total = int(lib.iiv_stream_length(n_emit, 1))
    out = torch.empty((total,), dtype=torch.uint8, device="cuda")
    bad = torch.zeros((1,), dtype=torch.int32, device="cuda")
    d_table = torch.from_numpy(np.ascontiguousarray(table, dtype=np.uint16).view(np.int16)).cuda()
    check(lib.iiv_emit_stream(
        m, opcode_records.data_ptr() if n else None, ticks.data_ptr() if n else None, n_emit,
        d_table.data_ptr(), ack, terminate, out.data_ptr(), total, bad.data_ptr(),
        torch.cuda.current_stream().cuda_stream))
    if int(bad.item()):
        raise KeyError("tick/page without a player opcode (ticks 4..66 even, pages 32..63)")
    return out


def stream_schedule(mode: str, n_frames: int, opcodes_per_frame: int = 980
                    ) -> List[Tuple[int, int, int]]:
    """Fixed-rate (frame, is_aux, budget) segments for synthetic workloads: exactly
    ``opcodes_per_frame`` pulls per frame and, in DHGR, a new generator per bank flip -- after
    291 tick opcodes in the first 2 KiB frame (7 header bytes), every 292 afterwards.
    ``plan_movie`` is the real thing (there the first frame is one tick short, because tick
    numbering starts at 1 while frame k is due at tick k * ticks_per_frame)."""
    segs = []
    aux = False
    count = 0
    next_flip = 291
    for fr in range(n_frames):
        left = opcodes_per_frame
        while left > 0:
            take = min(left, next_flip - count) if mode == "DHGR" else left
            segs.append((fr, int(aux), take))
            left -= take
            count += take
            if mode == "DHGR" and count == next_flip:
                aux = not aux
                next_flip += 292
    return segs

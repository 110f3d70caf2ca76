"""Encode a sequence of images as an optimized stream of screen changes
(reference transcoder/video.py), with the whole scoring + greedy emission loop of
an ``encode_frame`` call running in one CUDA kernel (``iiv_encode_clips``).

Drop-in surface: ``Video(frame_grabber, ticks_per_second, mode, palette)``,
``tick``, ``encode_frame(target, is_aux)`` yielding ``(page + 32, content,
[o0, o1, o2, o3])`` tuples that are bit-identical to the reference's for the same
frames, tables and ``random.seed`` / ``np.random.seed``.

How a lazy Python generator maps onto a kernel (SURVEY.md 8(b) "laziness"): the
reference runs one loop iteration per ``next()`` and ``Movie.encode`` abandons the
generator at every bank flip / new frame.  Here the first ``next()`` snapshots the
encoder state and both global MT19937 generators, runs the kernel speculatively for
``speculate`` opcodes and serves tuples from the result; pulling past the speculation
re-runs from the snapshot with a doubled budget (the kernel is deterministic, so the
prefix is identical).  When the generator is closed, abandoned or exhausted after k
pulls, the state is committed as of exactly k opcodes (one more launch from the
snapshot unless k happens to equal the speculated budget), the host-visible arrays
(``memory_map``, ``pixelmap.packed``, ``update_priority`` ...) are refreshed and the
global ``random`` / ``np.random`` generators are advanced to where the reference
would have left them.  Between those commit points the host arrays and global RNGs
lag behind; call ``Video.sync()`` to force a commit mid-generator.

For throughput (many clips, known schedules) use ``ops.encode_clips`` directly.
"""

import random
from typing import Iterator, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from . import screen
from .palette import Palette
from .video_mode import VideoMode

MAX_BUDGET = 2048   # kMaxBudget of csrc/iiv_encoder.cu


class Video:
    """Encodes sequence of images into prioritized screen byte changes."""

    CLOCK_SPEED = 1024 * 1024  # type: int

    def __init__(self, frame_grabber, ticks_per_second: float,
                 mode: VideoMode = VideoMode.HGR, palette: Palette = Palette.NTSC,
                 speculate: Optional[int] = None):
        self.mode = mode
        self.frame_grabber = frame_grabber
        self.ticks_per_second = float(ticks_per_second)
        self.ticks_per_frame = self.ticks_per_second / frame_grabber.input_frame_rate
        self.frame_number = 0
        self.palette = palette
        self._mode_name = "DHGR" if mode == VideoMode.DHGR else "HGR"

        # Initialize empty screen
        self.memory_map = screen.MemoryMap(screen_page=1)
        if self.mode == VideoMode.DHGR:
            self.aux_memory_map = screen.MemoryMap(screen_page=1)
            self.pixelmap = screen.DHGRBitmap(
                palette=palette, main_memory=self.memory_map,
                aux_memory=self.aux_memory_map)
        else:
            self.pixelmap = screen.HGRBitmap(palette=palette, main_memory=self.memory_map)

        # Accumulates pending edit weights across frames
        self.update_priority = np.zeros((32, 256), dtype=np.int32)
        if self.mode == VideoMode.DHGR:
            self.aux_update_priority = np.zeros((32, 256), dtype=np.int32)

        # Key is True for aux bank and False for main bank
        self.out_of_work = {True: False, False: False}

        # Movie.encode restarts the generator every 292 opcodes in DHGR (2 KiB of
        # stream) and every 980 at most in HGR (movie.py:94-102)
        self.speculate = int(speculate or (292 if mode == VideoMode.DHGR else 980))
        self._state = ops.new_clip_states(1)
        self._live = None   # the _Run of the generator currently being pulled

    def tick(self, ticks: int) -> bool:
        """Keep track of when it is time for a new image frame."""
        if ticks >= (self.ticks_per_frame * self.frame_number):
            self.frame_number += 1
            return True
        return False

    # -- host <-> device state ---------------------------------------------------------
    def _field(self, f, dtype, shape):
        return ops.state_field(self._state, f, dtype, shape)[0]

    def _upload(self):
        st = self._state
        st.zero_()
        self._field(ops.F_PACKED, torch.int64, (32, 128)).copy_(
            torch.from_numpy(self.pixelmap.packed.view(np.int64)))
        self._field(ops.F_MAIN, torch.uint8, (32, 256)).copy_(
            torch.from_numpy(self.memory_map.page_offset))
        self._field(ops.F_PRIO_MAIN, torch.int32, (32, 256)).copy_(
            torch.from_numpy(self.update_priority))
        if self.mode == VideoMode.DHGR:
            self._field(ops.F_AUX, torch.uint8, (32, 256)).copy_(
                torch.from_numpy(self.aux_memory_map.page_offset))
            self._field(ops.F_PRIO_AUX, torch.int32, (32, 256)).copy_(
                torch.from_numpy(self.aux_update_priority))
        mt = np.zeros(640, np.uint32)
        mt[:625] = ops.mt_from_numpy(np.random.get_state())
        self._field(ops.F_MT_NP, torch.int32, (640,)).copy_(torch.from_numpy(mt.view(np.int32)))
        mt[:625] = ops.mt_from_python(random.getstate())
        self._field(ops.F_MT_PY, torch.int32, (640,)).copy_(torch.from_numpy(mt.view(np.int32)))

    def _download(self, state: torch.Tensor):
        host = state.cpu()

        def field(f, dtype, shape):
            return ops.state_field(host, f, dtype, shape)[0].numpy()
        self.pixelmap.packed[...] = field(ops.F_PACKED, torch.int64, (32, 128)).view(np.uint64)
        self.memory_map.page_offset[...] = field(ops.F_MAIN, torch.uint8, (32, 256))
        self.update_priority[...] = field(ops.F_PRIO_MAIN, torch.int32, (32, 256))
        if self.mode == VideoMode.DHGR:
            self.aux_memory_map.page_offset[...] = field(ops.F_AUX, torch.uint8, (32, 256))
            self.aux_update_priority[...] = field(ops.F_PRIO_AUX, torch.int32, (32, 256))
        mt_np = field(ops.F_MT_NP, torch.int32, (640,)).view(np.uint32)[:625]
        mt_py = field(ops.F_MT_PY, torch.int32, (640,)).view(np.uint32)[:625]
        np.random.set_state(ops.mt_to_numpy(mt_np))
        random.setstate(ops.mt_to_python(mt_py))
        flags = int(field(ops.F_FLAGS, torch.int32, (8,))[2])
        if flags & 1:
            raise AssertionError("DHGR content byte with bit 7 set")   # video.py:135-137
        if flags & ~1:
            raise RuntimeError("encoder kernel internal error %#x" % flags)

    # -- encode_frame ------------------------------------------------------------------------
    def encode_frame(self, target: screen.Bitmap, is_aux: bool
                     ) -> Iterator[Tuple[int, int, List[int]]]:
        """Converge towards target frame in priority order of edit distance."""
        if self._live is not None:      # previous generator abandoned without close()
            self._live.commit()
        if is_aux:
            memory_map, update_priority = self.aux_memory_map, self.aux_update_priority
        else:
            memory_map, update_priority = self.memory_map, self.update_priority

        # Make sure nothing is leaking into screen holes
        assert np.count_nonzero(memory_map.page_offset[screen.SCREEN_HOLES]) == 0
        print("Similarity %f" % (update_priority.mean()))

        run = _Run(self, target, bool(is_aux))
        self._live = run
        try:
            k = 0
            while True:
                op = run.opcode(k)
                k += 1
                run.pulled = k      # the stores of opcode k-1 are committed state
                yield op
        finally:
            run.commit()

    def sync(self) -> None:
        """Commit the live generator's state as of the opcodes pulled so far."""
        if self._live is not None:
            self._live.commit(keep=True)


class _Run:
    """One encode_frame generator: snapshot, speculative kernel runs, commit."""

    def __init__(self, video: Video, target: screen.Bitmap, is_aux: bool):
        self.v = video
        self.is_aux = is_aux
        self.pulled = 0
        self.closed = False
        m = video._mode_name
        video._upload()
        self.snapshot = video._state.clone()
        banks = [target.main_memory.page_offset]
        if m == "DHGR":
            banks.append(target.aux_memory.page_offset)
        self.tmem = torch.from_numpy(np.ascontiguousarray(np.stack(banks))).cuda().view(
            1, 1, len(banks), 32, 256)
        self.tpacked = torch.from_numpy(
            np.ascontiguousarray(target.packed).view(np.int64)).cuda().view(1, 1, 32, 128)
        self.table = type(target).edit_distances_device(target.palette)
        self.pad = (32, int(banks[1 if (m == "DHGR" and is_aux) else 0][0, 0]), [0, 0, 0, 0])
        self.budget = 0
        self.ops = None
        self.real = 0
        self.state_after = None
        self._run(min(video.speculate, MAX_BUDGET))

    def _run(self, budget: int):
        state = self.snapshot.clone()
        opc, info = ops.encode_clips(self.v._mode_name, state, self.tmem, self.tpacked,
                                     [(0, int(self.is_aux), budget)], self.table)
        torch.cuda.current_stream().synchronize()
        self.ops = opc.cpu().numpy()[0]
        self.real = int(info.cpu().numpy()[0, 0, 0])
        self.budget = budget
        self.state_after = state

    def opcode(self, k: int):
        if k >= self.budget and self.real == self.budget:
            if self.budget >= MAX_BUDGET:
                raise NotImplementedError(
                    "more than %d opcodes pulled from one encode_frame generator" % MAX_BUDGET)
            self._run(min(2 * self.budget, MAX_BUDGET))
        if k >= self.real:
            # out of work: (32, target[0, 0], [0, 0, 0, 0]) forever (video.py:249-251)
            self.v.out_of_work[self.is_aux] = True
            return self.pad
        r = self.ops[k]
        return int(r[0]), int(r[1]), [int(r[2]), int(r[3]), int(r[4]), int(r[5])]

    def commit(self, keep: bool = False):
        if self.closed:
            return
        k = self.pulled
        if k > 0:
            if self.real < self.budget and k >= self.real:
                pass                      # ran dry: state_after is final whatever k is
            elif k != self.budget:
                self._run(k)
            self.v._state.copy_(self.state_after)
            self.v._download(self.state_after)
        if not keep:
            self.closed = True
            if self.v._live is self:
                self.v._live = None

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

The speculated budget is a guess at how many opcodes the caller will pull before it
abandons the generator; ``Movie.encode`` does so at every bank flip (a fixed number of
opcodes apart) and at every new frame, so the guess is learnt from the generators seen
so far (``PullPredictor``).  A wrong guess costs one more launch, never a different
result.  Per generator the host side moves one state blob up and one down through
page-locked staging buffers; the launch itself allocates nothing.

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


class PullPredictor:
    """Guesses how many opcodes the next ``encode_frame`` generator will be asked for.

    ``Movie.encode`` (movie.py:94-102) drops a generator at every bank flip, a fixed number
    of opcodes after the previous one, and at every new frame.  Both periods are taken from
    what the caller did so far; until they are known the guess is ``default``.  A wrong
    guess costs one more kernel launch, never a different result."""

    def __init__(self, default: int, flips: bool, limit: int):
        self.default, self.flips, self.limit = int(default), bool(flips), int(limit)
        self._target = None
        self._aux = None
        self._since_flip = 0       # opcodes pulled since the bank last changed
        self._since_frame = 0      # ... since the target last changed
        self._flip_period = None   # opcodes between the last two bank changes
        self._frame_period = None  # opcodes pulled for the last complete target

    def guess(self, target, is_aux: bool) -> int:
        """To be called once per generator, before its first pull."""
        if target is not self._target and self._target is not None:
            self._frame_period = self._since_frame
            self._since_frame = 0
        if self._aux is not None and bool(is_aux) != self._aux:
            self._flip_period = self._since_flip
            self._since_flip = 0
        self._target, self._aux = target, bool(is_aux)
        n = self.default
        if self.flips:
            period = self._flip_period or self.default
            if period > self._since_flip:
                n = period - self._since_flip
        if self._frame_period and self._frame_period > self._since_frame:
            n = min(n, self._frame_period - self._since_frame)
        return max(1, min(n, self.limit))

    def pulled(self, n: int) -> None:
        """The generator last guessed for was pulled ``n`` times in all."""
        self._since_flip += n
        self._since_frame += n


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
        self._adaptive = speculate is None
        self._state = ops.new_clip_states(1)
        self._live = None   # the _Run of the generator currently being pulled
        # page-locked staging: the state blob on its way up / down, opcode records and
        # segment info on their way down
        self._stage_up = torch.zeros(ops.STATE_BYTES, dtype=torch.uint8, pin_memory=True)
        self._stage_down = torch.zeros(ops.STATE_BYTES, dtype=torch.uint8, pin_memory=True)
        self._stage_ops = torch.zeros(MAX_BUDGET * 8, dtype=torch.uint8, pin_memory=True)
        self._stage_info = torch.zeros(8, dtype=torch.int64, pin_memory=True)
        self._d_ops = torch.empty((1, MAX_BUDGET, 8), dtype=torch.uint8, device="cuda")
        self._d_info = torch.zeros((1, 1, 8), dtype=torch.int64, device="cuda")
        self._plans = {}
        self._predictor = PullPredictor(self.speculate, mode == VideoMode.DHGR, MAX_BUDGET)

    def tick(self, ticks: int) -> bool:
        """Keep track of when it is time for a new image frame."""
        if ticks >= (self.ticks_per_frame * self.frame_number):
            self.frame_number += 1
            return True
        return False

    # -- host <-> device state ---------------------------------------------------------
    @staticmethod
    def _host_field(blob: np.ndarray, f, dtype, shape) -> np.ndarray:
        off = ops.STATE_OFFSETS[f]
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return blob[off:off + n].view(dtype).reshape(shape)

    def _upload(self):
        """Host arrays and the process-global generators -> the device state blob."""
        blob = self._stage_up.numpy()
        blob[:] = 0
        hf = self._host_field
        hf(blob, ops.F_PACKED, np.uint64, (32, 128))[...] = self.pixelmap.packed
        hf(blob, ops.F_MAIN, np.uint8, (32, 256))[...] = self.memory_map.page_offset
        hf(blob, ops.F_PRIO_MAIN, np.int32, (32, 256))[...] = self.update_priority
        if self.mode == VideoMode.DHGR:
            hf(blob, ops.F_AUX, np.uint8, (32, 256))[...] = self.aux_memory_map.page_offset
            hf(blob, ops.F_PRIO_AUX, np.int32, (32, 256))[...] = self.aux_update_priority
        hf(blob, ops.F_MT_NP, np.uint32, (625,))[...] = ops.mt_from_numpy(np.random.get_state())
        hf(blob, ops.F_MT_PY, np.uint32, (625,))[...] = ops.mt_from_python(random.getstate())
        self._state[0].copy_(self._stage_up, non_blocking=True)

    def _download(self, blob: np.ndarray):
        """A state blob brought back by _Run._run -> host arrays and global generators."""
        hf = self._host_field
        self.pixelmap.packed[...] = hf(blob, ops.F_PACKED, np.uint64, (32, 128))
        self.memory_map.page_offset[...] = hf(blob, ops.F_MAIN, np.uint8, (32, 256))
        self.update_priority[...] = hf(blob, ops.F_PRIO_MAIN, np.int32, (32, 256))
        if self.mode == VideoMode.DHGR:
            self.aux_memory_map.page_offset[...] = hf(blob, ops.F_AUX, np.uint8, (32, 256))
            self.aux_update_priority[...] = hf(blob, ops.F_PRIO_AUX, np.int32, (32, 256))
        np.random.set_state(ops.mt_to_numpy(hf(blob, ops.F_MT_NP, np.uint32, (625,))))
        random.setstate(ops.mt_to_python(hf(blob, ops.F_MT_PY, np.uint32, (625,))))
        flags = int(hf(blob, ops.F_FLAGS, np.int32, (8,))[2])
        if flags & 1:
            raise AssertionError("DHGR content byte with bit 7 set")   # video.py:135-137
        if flags & ~1:
            raise RuntimeError("encoder kernel internal error %#x" % flags)

    def _plan(self, is_aux: bool, budget: int) -> ops.SegmentPlan:
        key = (bool(is_aux), int(budget))
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) > 4096:
                self._plans.clear()
            plan = self._plans[key] = ops.SegmentPlan([(0, int(is_aux), int(budget))])
        return plan

    def _predict(self, target, is_aux: bool) -> int:
        if not self._adaptive:
            return self.speculate
        return self._predictor.guess(target, is_aux)

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
        guess = video._predict(target, is_aux)
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
        self.state_after = None      # host copy of the state blob after `budget` opcodes
        self._run(guess)

    def _run(self, budget: int):
        v = self.v
        state = self.snapshot.clone()
        d_ops = v._d_ops[:, :budget]
        ops.encode_clips(v._mode_name, state, self.tmem, self.tpacked,
                         v._plan(self.is_aux, budget), self.table,
                         opcodes=d_ops, seg_info=v._d_info)
        v._stage_ops[:budget * 8].copy_(d_ops.reshape(-1), non_blocking=True)
        v._stage_info.copy_(v._d_info.view(-1), non_blocking=True)
        v._stage_down.copy_(state[0], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.ops = v._stage_ops[:budget * 8].numpy().reshape(budget, 8).tolist()
        self.real = int(v._stage_info[0])
        self.budget = budget
        self.state_after = v._stage_down.numpy().copy()

    def opcode(self, k: int):
        if k >= self.budget and self.real == self.budget:
            if self.budget >= MAX_BUDGET:
                raise NotImplementedError(
                    "more than %d opcodes pulled from one encode_frame generator" % MAX_BUDGET)
            self._run(min(max(2 * self.budget, self.v.speculate), MAX_BUDGET))
        if k >= self.real:
            # out of work: (32, target[0, 0], [0, 0, 0, 0]) forever (video.py:249-251)
            self.v.out_of_work[self.is_aux] = True
            return self.pad
        r = self.ops[k]
        return r[0], r[1], r[2:6]

    def commit(self, keep: bool = False):
        if self.closed:
            return
        k = self.pulled
        if k > 0:
            if self.real < self.budget and k >= self.real:
                pass                      # ran dry: state_after is final whatever k is
            elif k != self.budget:
                self._run(k)
            self.v._download(self.state_after)
        if not keep:
            self.closed = True
            self.v._predictor.pulled(k)
            if self.v._live is self:
                self.v._live = None

"""Encode a sequence of images as an optimized stream of screen changes
(reference transcoder/video.py), with the whole scoring + greedy emission loop of
an ``encode_frame`` call running in one CUDA kernel (``iiv_encode_clips``).

Drop-in surface: ``Video(frame_grabber, ticks_per_second, mode, palette)``,
``tick``, ``encode_frame(target, is_aux)`` yielding ``(page + 32, content,
[o0, o1, o2, o3])`` tuples that are bit-identical to the reference's for the same
frames, tables and ``random.seed`` / ``np.random.seed``.

How a lazy Python generator maps onto a kernel (SURVEY.md 8(b) "laziness"): the
reference runs one loop iteration per ``next()`` and ``Movie.encode`` abandons the
generator at every bank flip / new frame.  Here the first ``next()`` runs the kernel
for a predicted number of opcodes from the committed encoder state and serves tuples
from the result; pulling past the prediction re-runs from the same state with a doubled
budget (the kernel is deterministic, so the prefix is identical).  When the generator is
closed, abandoned or exhausted after k pulls, the state as of exactly k opcodes becomes
the committed one (one more launch unless k equals the predicted budget) and the global
``random`` / ``np.random`` generators are advanced by exactly the words the reference
would have drawn.

The encoder state lives on the device.  ``memory_map``, ``aux_memory_map``,
``pixelmap``, ``update_priority`` and ``aux_update_priority`` are the reference's
attributes, brought up to date when they are read (mid-generator: as of the opcodes
pulled so far) and sent back to the device before the next generator if nothing was
pulled in between -- so edits made between generators are honoured, while a caller that
never looks (``Movie.encode``) pays for no copies.  Edits made to them in the middle of
a generator are not seen by that generator.

Launches are pipelined: ``Movie.encode`` drops a generator at every bank flip, a fixed
number of opcodes apart, so while the host hands out the opcodes of one generator the
kernel of the predicted next one (same target, other bank) is already running from the
state the current one will leave.  A wrong prediction costs the speculated launch, never
a different result.  Other code drawing from the global ``random`` / ``np.random``
between two generators is detected (``rng_check``) and the device copies of the
generators are refreshed.

For throughput (many clips, known schedules) use ``ops.encode_clips`` directly.
"""

import array
import ctypes
import random
import sys
import weakref
from typing import Iterator, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from . import screen
from ._lib import check, lib
from .palette import Palette
from .video_mode import VideoMode

MAX_BUDGET = 1 << 17   # kMaxBudget of csrc/iiv_encoder.cu: opcodes per generator
_STAGE_OPCODES = 2048   # opcode records a regular staging buffer holds


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

    def peek_next(self, n: int) -> Optional[int]:
        """If the generator last guessed for is pulled exactly ``n`` times and then dropped
        for a bank flip, how many opcodes will its successor (same target, other bank) be
        asked for?  None when that cannot be told yet or a new frame is due instead."""
        if not self.flips or not self._flip_period or not self._frame_period:
            return None
        if self._since_flip + n != self._flip_period:
            return None                     # it will not end in a flip
        left = self._frame_period - (self._since_frame + n)
        if left <= 0:
            return None                     # a new target comes next
        return max(1, min(self._flip_period, left, self.limit))


class _GlobalRng:
    """The two process-global MT19937 generators the reference draws from: ``np.random``
    (video.py:265) and ``random`` (video.py:178, :291).  numpy's state is reached in place
    through the bit generator's ctypes interface (a 2.5 KB read instead of a 55 us
    ``get_state``); CPython's only through ``getstate``."""

    def __init__(self):
        self._np_view = None
        try:
            bg = np.random.mtrand._rand._bit_generator
            addr = int(bg.ctypes.state_address)
            view = np.ctypeslib.as_array(
                ctypes.cast(addr, ctypes.POINTER(ctypes.c_uint32)), shape=(625,))
            st = np.random.get_state()
            if st[0] == "MT19937" and np.array_equal(view[:624], st[1]) and int(view[624]) == st[2]:
                self._np_view, self._np_bitgen = view, bg
        except Exception:   # noqa: BLE001  (layout not as expected: use the public API)
            self._np_view = None

    def numpy_words(self) -> np.ndarray:
        """uint32[625]: key + position of np.random's generator."""
        if self._np_view is not None:
            return self._np_view.copy()
        return ops.mt_from_numpy(np.random.get_state())

    @staticmethod
    def python_words() -> np.ndarray:
        """uint32[625]: state + index of the random module's generator."""
        version, internal, _ = random.getstate()
        if version != 3 or len(internal) != 625:
            raise ValueError("unexpected random.getstate() layout")
        return np.frombuffer(array.array("I", internal), dtype=np.uint32)

    def numpy_mark(self) -> bytes:
        if self._np_view is not None:
            return self._np_view.tobytes()
        return self.numpy_words().tobytes()

    @staticmethod
    def python_mark():
        return random.getstate()[1]

    def set_states(self, numpy_words: np.ndarray, python_words: np.ndarray, gauss_next):
        """Put both generators where the device copies stand (uint32[625] each: state +
        position).  Returns the python state tuple, which doubles as its mark.  Cached
        gaussians are left alone: numpy's live outside the bit generator, CPython's is
        passed back in."""
        if self._np_view is not None:
            self._np_view[:] = numpy_words
        else:
            st = np.random.get_state()
            np.random.set_state(("MT19937", np.array(numpy_words[:624], dtype=np.uint32),
                                 int(numpy_words[624])) + tuple(st[3:]))
        internal = tuple(python_words.tolist())
        random.setstate((3, internal, gauss_next))
        return internal


class _Stage:
    """Page-locked landing buffers for one kernel run: opcode records, segment info and
    the tail of the state blob (both generators and the flags)."""

    def __init__(self, n_opcodes: int, tail_bytes: int):
        self.capacity = n_opcodes
        self.ops = torch.zeros(n_opcodes * 8, dtype=torch.uint8, pin_memory=True)
        self.info = torch.zeros(8, dtype=torch.int64, pin_memory=True)
        self.tail = torch.zeros(tail_bytes, dtype=torch.uint8, pin_memory=True)


class _Kernel:
    """One launch: ``budget`` opcodes of (target, bank) from the state in slot ``base``,
    leaving the state in slot ``out``."""

    __slots__ = ("tgt", "is_aux", "budget", "base", "out", "stage", "event", "_ops",
                 "real", "similarity", "words_np", "words_py", "flags", "mt_np", "mt_py")

    def __init__(self, tgt, is_aux, budget, base, out, stage, event):
        self.tgt, self.is_aux, self.budget = tgt, is_aux, budget
        self.base, self.out, self.stage, self.event = base, out, stage, event
        self._ops = None

    def wait(self, tail_flags_offset: int, events: list):
        if self._ops is None:
            check(lib.iiv_event_wait(self.event))
            events.append(self.event)          # waited for: the handle can be recorded again
            self.event = None
            n = self.budget
            self._ops = self.stage.ops[:n * 8].numpy().reshape(n, 8).tolist()
            info = self.stage.info.numpy()
            self.real = int(info[0])
            self.similarity = float(info[1]) / 8192.0
            self.words_np, self.words_py = int(info[2]), int(info[3])
            tail = self.stage.tail.numpy()
            self.flags = int(tail[tail_flags_offset:tail_flags_offset + 32].view(np.int32)[2])
            # both generators after `budget` opcodes (uint32[640] slots: 624 words + position)
            self.mt_np = tail[:2500].view(np.uint32).copy()
            self.mt_py = tail[2560:2560 + 2500].view(np.uint32).copy()
        return self._ops


class _Target:
    __slots__ = ("ref", "main", "aux", "packed", "tmem", "tpacked", "table", "from_bitmap",
                 "pad")


class Video:
    """Encodes sequence of images into prioritized screen byte changes."""

    CLOCK_SPEED = 1024 * 1024  # type: int
    _N_SLOTS = 4

    def __init__(self, frame_grabber, ticks_per_second: float,
                 mode: VideoMode = VideoMode.HGR, palette: Palette = Palette.NTSC,
                 speculate: Optional[int] = None, rng_check: bool = True,
                 pipeline: bool = True):
        self.mode = mode
        self.frame_grabber = frame_grabber
        self.ticks_per_second = float(ticks_per_second)
        self.ticks_per_frame = self.ticks_per_second / frame_grabber.input_frame_rate
        self.frame_number = 0
        self.palette = palette
        self._mode_name = "DHGR" if mode == VideoMode.DHGR else "HGR"
        self._dhgr = mode == VideoMode.DHGR

        # Initialize empty screen
        self._memory_map = screen.MemoryMap(screen_page=1)
        if self._dhgr:
            self._aux_memory_map = screen.MemoryMap(screen_page=1)
            self._pixelmap = screen.DHGRBitmap(
                palette=palette, main_memory=self._memory_map,
                aux_memory=self._aux_memory_map)
        else:
            self._aux_memory_map = None
            self._pixelmap = screen.HGRBitmap(palette=palette, main_memory=self._memory_map)

        # Accumulates pending edit weights across frames
        self._update_priority = np.zeros((32, 256), dtype=np.int32)
        self._aux_update_priority = np.zeros((32, 256), dtype=np.int32) if self._dhgr else None

        # Key is True for aux bank and False for main bank
        self.out_of_work = {True: False, False: False}

        # Movie.encode restarts the generator every 292 opcodes in DHGR (2 KiB of
        # stream) and every 980 at most in HGR (movie.py:94-102)
        self.speculate = int(speculate or (292 if self._dhgr else 980))
        self._adaptive = speculate is None
        self._pipeline = bool(pipeline) and self._dhgr
        self._rng_check = bool(rng_check)
        self._predictor = PullPredictor(self.speculate, self._dhgr, MAX_BUDGET)
        self._rng = _GlobalRng()

        # device side: state slots (committed state, runs in flight), landing buffers
        self._slots = ops.new_clip_states(self._N_SLOTS)
        self._committed = 0
        self._free_slots = list(range(1, self._N_SLOTS))
        self._arrays_end = ops.STATE_OFFSETS[ops.F_MT_NP]      # packed .. priorities
        self._tail_bytes = ops.STATE_BYTES - self._arrays_end   # both generators + flags
        self._flags_off = ops.STATE_OFFSETS[ops.F_FLAGS] - self._arrays_end
        self._stage_up = torch.zeros(ops.STATE_BYTES, dtype=torch.uint8, pin_memory=True)
        self._stage_down = torch.zeros(ops.STATE_BYTES, dtype=torch.uint8, pin_memory=True)
        self._stages = [_Stage(_STAGE_OPCODES, self._tail_bytes) for _ in range(4)]
        self._stage_next = 0
        self._d_ops = torch.empty(_STAGE_OPCODES * 8, dtype=torch.uint8, device="cuda")
        self._targets = {}
        self._events = []           # cudaEvent_t handles free for the next launch
        self._mode_id = ops.mode_id(self._mode_name)
        self._stream = torch.cuda.current_stream().cuda_stream

        self._live = None           # the _Run of the generator currently being pulled
        self._spec = None           # _Kernel launched ahead for the predicted next generator
        self._pulls_closed = 0      # opcodes pulled from the generators that are closed
        self._host_fresh = True     # host arrays == committed state
        self._touched_at = 0        # value of _pulls when the host arrays were last handed out
        self._touched = True        # ... and they may have been edited since (upload first)
        self._np_mark = None        # global generator states the device copies correspond to
        self._py_mark = None
        self._gauss_next = None

    @property
    def _pulls(self) -> int:
        """Opcodes pulled over this object's life."""
        live = self._live
        return self._pulls_closed + (live.pulled if live is not None else 0)

    def __del__(self):
        if sys is None or sys.is_finalizing():
            return
        try:
            for ev in self._events:
                lib.iiv_event_destroy(ev)
        except Exception:   # noqa: BLE001
            pass

    # -- the reference's attributes, synchronised on access --------------------------------
    def _touch(self):
        self._sync_host()
        self._touched, self._touched_at = True, self._pulls

    @property
    def memory_map(self):
        self._touch()
        return self._memory_map

    @memory_map.setter
    def memory_map(self, value):
        self._touch()
        self._memory_map = value

    @property
    def aux_memory_map(self):
        if not self._dhgr:      # the reference's HGR Video has no such attribute (video.py:44-53)
            raise AttributeError("aux_memory_map")
        self._touch()
        return self._aux_memory_map

    @aux_memory_map.setter
    def aux_memory_map(self, value):
        self._touch()
        self._aux_memory_map = value

    @property
    def pixelmap(self):
        self._touch()
        return self._pixelmap

    @pixelmap.setter
    def pixelmap(self, value):
        self._touch()
        self._pixelmap = value

    @property
    def update_priority(self):
        self._touch()
        return self._update_priority

    @update_priority.setter
    def update_priority(self, value):
        self._touch()
        self._update_priority = value

    @property
    def aux_update_priority(self):
        if not self._dhgr:
            raise AttributeError("aux_update_priority")
        self._touch()
        return self._aux_update_priority

    @aux_update_priority.setter
    def aux_update_priority(self, value):
        self._touch()
        self._aux_update_priority = value

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

    def _upload_arrays(self):
        """Host arrays -> the committed state slot (they may have been edited)."""
        mm = self._memory_map.page_offset
        # Make sure nothing is leaking into screen holes (video.py:87-88)
        assert np.count_nonzero(mm[screen.SCREEN_HOLES]) == 0
        blob = self._stage_up.numpy()
        hf = self._host_field
        hf(blob, ops.F_PACKED, np.uint64, (32, 128))[...] = self._pixelmap.packed
        hf(blob, ops.F_MAIN, np.uint8, (32, 256))[...] = mm
        hf(blob, ops.F_PRIO_MAIN, np.int32, (32, 256))[...] = self._update_priority
        if self._dhgr:
            am = self._aux_memory_map.page_offset
            assert np.count_nonzero(am[screen.SCREEN_HOLES]) == 0
            hf(blob, ops.F_AUX, np.uint8, (32, 256))[...] = am
            hf(blob, ops.F_PRIO_AUX, np.int32, (32, 256))[...] = self._aux_update_priority
        n = self._arrays_end
        self._slots[self._committed, :n].copy_(self._stage_up[:n], non_blocking=True)
        # the staging buffer is reused: the copy must have left it before the next fill
        torch.cuda.current_stream().synchronize()

    def _upload_rng(self):
        """Both process-global generators -> the committed state slot."""
        blob = self._stage_up.numpy()
        hf = self._host_field
        hf(blob, ops.F_MT_NP, np.uint32, (625,))[...] = self._rng.numpy_words()
        hf(blob, ops.F_MT_PY, np.uint32, (625,))[...] = self._rng.python_words()
        a, b = self._arrays_end, ops.STATE_OFFSETS[ops.F_FLAGS]
        self._slots[self._committed, a:b].copy_(self._stage_up[a:b], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._np_mark = self._rng.numpy_mark()
        self._py_mark = self._rng.python_mark()

    def _download_slot(self, slot: int):
        """A state slot -> the host arrays (in place: the Bitmap aliases the memory maps)."""
        n = self._arrays_end
        self._stage_down[:n].copy_(self._slots[slot, :n], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        blob = self._stage_down.numpy()
        hf = self._host_field
        self._pixelmap.packed[...] = hf(blob, ops.F_PACKED, np.uint64, (32, 128))
        self._memory_map.page_offset[...] = hf(blob, ops.F_MAIN, np.uint8, (32, 256))
        self._update_priority[...] = hf(blob, ops.F_PRIO_MAIN, np.int32, (32, 256))
        if self._dhgr:
            self._aux_memory_map.page_offset[...] = hf(blob, ops.F_AUX, np.uint8, (32, 256))
            self._aux_update_priority[...] = hf(blob, ops.F_PRIO_AUX, np.int32, (32, 256))

    def _sync_host(self):
        """Bring the host arrays up to date: the committed state, or -- in the middle of a
        generator -- the state as of the opcodes pulled so far."""
        if self._live is not None:
            self._live.sync_host()
        elif not self._host_fresh:
            self._download_slot(self._committed)
            self._host_fresh = True

    def sync(self) -> None:
        """Make the host arrays current (reading any of them does the same)."""
        self._sync_host()

    # -- launches ---------------------------------------------------------------------------
    def _target(self, target) -> _Target:
        """Device copies of a target bitmap, uploaded once per target (per frame) and checked
        against the host arrays every time (they are the reference's inputs)."""
        banks = [target.main_memory.page_offset]
        if self._dhgr:
            banks.append(target.aux_memory.page_offset)
        # the bitmap's own device copies (screen.Bitmap._pack) serve as long as its host
        # arrays still hold what they were made from; its packed words only need checking
        # once they have been brought to the host, where somebody might have edited them
        dev = getattr(target, "_device_copy", None)
        dev_ok = (dev is not None and dev[0].shape[0] == len(banks)
                  and np.array_equal(dev[0][0], banks[0])
                  and (not self._dhgr or np.array_equal(dev[0][1], banks[1]))
                  and (getattr(target, "_packed", None) is None
                       or (dev[1] is not None and np.array_equal(dev[1], target._packed))))
        ent = self._targets.get(id(target))
        if ent is not None and ent.ref() is target:
            if ent.from_bitmap and dev_ok and ent.tpacked.data_ptr() == dev[3].data_ptr():
                return ent
            if (not ent.from_bitmap and np.array_equal(ent.packed, target.packed)
                    and np.array_equal(ent.main, banks[0])
                    and (not self._dhgr or np.array_equal(ent.aux, banks[1]))):
                return ent
        ent = _Target()
        ent.ref = weakref.ref(target)
        ent.from_bitmap = dev_ok
        if dev_ok:
            ent.main = ent.aux = ent.packed = None
            ent.tmem = dev[2].view(1, 1, len(banks), 32, 256)
            ent.tpacked = dev[3].view(1, 1, 32, 128)
        else:
            ent.main = banks[0].copy()
            ent.aux = banks[1].copy() if self._dhgr else None
            ent.packed = np.array(target.packed, dtype=np.uint64)
            ent.tmem = torch.from_numpy(np.ascontiguousarray(np.stack(banks))).cuda().view(
                1, 1, len(banks), 32, 256)
            ent.tpacked = torch.from_numpy(ent.packed.view(np.int64)).cuda().view(1, 1, 32, 128)
        ent.table = type(target).edit_distances_device(target.palette)
        ent.pad = (int(banks[0][0, 0]), int(banks[1][0, 0]) if self._dhgr else 0)
        if len(self._targets) >= 4:
            self._targets.clear()
        self._targets[id(target)] = ent
        return ent

    def _take_slot(self, keep) -> int:
        for k, s in enumerate(self._free_slots):
            if s not in keep:
                return self._free_slots.pop(k)
        raise RuntimeError("no free encoder state slot")

    def _stage_for(self, budget: int) -> _Stage:
        if budget > _STAGE_OPCODES:
            return _Stage(budget, self._tail_bytes)
        st = self._stages[self._stage_next]
        self._stage_next = (self._stage_next + 1) % len(self._stages)
        return st

    def _launch(self, base: int, tgt: _Target, is_aux: bool, budget: int) -> _Kernel:
        """One library call, one kernel: start from slot ``base``, leave the state in a free
        slot, write opcodes, info and the state's tail straight into page-locked host
        memory, record an event."""
        out = self._take_slot(keep=(base,))
        stage = self._stage_for(budget)
        if budget * 8 > self._d_ops.numel():
            self._d_ops = torch.empty(budget * 8, dtype=torch.uint8, device="cuda")
        ev = self._events.pop() if self._events else lib.iiv_event_create()
        if not ev:
            raise RuntimeError("cudaEventCreate failed")
        slots = self._slots.data_ptr()
        check(lib.iiv_encode_generator(
            self._mode_id, slots + base * ops.STATE_BYTES, slots + out * ops.STATE_BYTES,
            tgt.tmem.data_ptr(), tgt.tpacked.data_ptr(), int(is_aux), int(budget),
            tgt.table.data_ptr(), self._d_ops.data_ptr(), stage.ops.data_ptr(),
            stage.info.data_ptr(), stage.tail.data_ptr(), ev, self._stream))
        return _Kernel(tgt, is_aux, budget, base, out, stage, ev)

    def _release(self, kernel: Optional[_Kernel], keep=()):
        """Give a finished or abandoned run's output slot back."""
        if kernel is None:
            return
        if kernel.event is not None:           # never waited for: re-recording it is fine
            self._events.append(kernel.event)
            kernel.event = None
        if kernel.out not in keep and kernel.out != self._committed \
                and kernel.out not in self._free_slots:
            self._free_slots.append(kernel.out)

    def _drop_spec(self):
        if self._spec is not None:
            self._release(self._spec)
            self._spec = None

    def _predict(self, target, is_aux: bool) -> int:
        if not self._adaptive:
            return self.speculate
        return self._predictor.guess(target, is_aux)

    # -- encode_frame ------------------------------------------------------------------------
    def encode_frame(self, target: screen.Bitmap, is_aux: bool
                     ) -> Iterator[Tuple[int, int, List[int]]]:
        """Converge towards target frame in priority order of edit distance."""
        if self._live is not None:      # previous generator abandoned without close()
            self._live.finish()
        run = _Run(self, target, bool(is_aux))
        self._live = run
        # (video.py:90; the mean is taken over the priorities the kernel scored from)
        print("Similarity %f" % run.kernel.similarity)
        try:
            k = 0
            while True:
                # the opcodes the current kernel run holds, without a call per pull
                kern = run.kernel
                rows = kern._ops
                n = min(kern.real, kern.budget)
                while k < n:
                    r = rows[k]
                    k += 1
                    run.pulled = k      # the stores of opcode k-1 are committed state
                    yield r[0], r[1], r[2:6]
                # past it: re-run with a larger budget, or out of work (padding)
                op = run.opcode(k)
                k += 1
                run.pulled = k
                yield op
        finally:
            run.finish()


class _Run:
    """One encode_frame generator: a kernel run from the committed state (possibly one
    launched ahead of time), re-runs when the caller pulls past it, commit at the end."""

    def __init__(self, video: Video, target: screen.Bitmap, is_aux: bool):
        v = self.v = video
        self.is_aux = is_aux
        self.pulled = 0
        self.closed = False
        self.synced_at = -1          # pulls at which the host arrays were last made current
        v._stream = torch.cuda.current_stream().cuda_stream
        guess = v._predict(target, is_aux)
        tgt = v._target(target)
        self.pad = (32, tgt.pad[1 if (v._dhgr and is_aux) else 0], [0, 0, 0, 0])
        # host-side edits made since the last pull go to the device first
        dirty = v._touched and v._touched_at == v._pulls
        v._touched = False
        py_state = random.getstate() if (v._rng_check or v._py_mark is None) else None
        self.gauss_next = py_state[2] if py_state is not None else v._gauss_next
        v._gauss_next = self.gauss_next
        rng_moved = (v._np_mark is None or v._rng.numpy_mark() != v._np_mark
                     or (py_state is not None and py_state[1] != v._py_mark))
        spec = v._spec
        v._spec = None
        if dirty or rng_moved or spec is None or spec.tgt is not tgt \
                or spec.is_aux != is_aux or spec.base != v._committed:
            if spec is not None:
                v._release(spec)
            if dirty:
                v._upload_arrays()
                v._host_fresh = True
            if rng_moved:
                v._upload_rng()
            self.kernel = v._launch(v._committed, tgt, is_aux, guess)
        else:
            self.kernel = spec           # already running (or done) from the right state
        self.base = v._committed
        self._speculate_next()
        self.kernel.wait(v._flags_off, v._events)

    def _speculate_next(self):
        """Launch the predicted successor (same target, other bank) behind this run."""
        v, k = self.v, self.kernel
        if not (v._pipeline and v._adaptive):
            return
        n = v._predictor.peek_next(k.budget)
        if n is None or len(v._free_slots) < 2:
            return
        v._spec = v._launch(k.out, k.tgt, not self.is_aux, n)

    def _rerun(self, budget: int):
        v = self.v
        v._drop_spec()                       # it was based on the run being replaced
        old = self.kernel
        self.kernel = v._launch(self.base, old.tgt, self.is_aux, budget)
        v._release(old)
        self.kernel.wait(v._flags_off, v._events)

    def opcode(self, k: int):
        kern = self.kernel
        if k >= kern.budget and kern.real == kern.budget:
            if k >= MAX_BUDGET:
                raise NotImplementedError(
                    "more than %d opcodes pulled from one encode_frame generator" % MAX_BUDGET)
            # (the current run may be shorter than k: sync() settles on exactly the opcodes
            # pulled while the generator goes on serving the run it started with)
            self._rerun(min(max(2 * kern.budget, self.v.speculate, k + 1), MAX_BUDGET))
            kern = self.kernel
        if k >= kern.real:
            # out of work: (32, target[0, 0], [0, 0, 0, 0]) forever (video.py:249-251)
            self.v.out_of_work[self.is_aux] = True
            return self.pad
        r = kern._ops[k]
        return r[0], r[1], r[2:6]

    def _settle(self) -> _Kernel:
        """The run whose final state is the state after exactly ``pulled`` opcodes."""
        kern, k = self.kernel, self.pulled
        if kern.real < kern.budget and k >= kern.real:
            return kern                       # ran dry: the state is final whatever k is
        if k != kern.budget:
            self._rerun(k)
        return self.kernel

    def _check(self, kern: _Kernel):
        if kern.flags & 1:
            raise AssertionError("DHGR content byte with bit 7 set")   # video.py:135-137
        if kern.flags & ~1:
            raise RuntimeError("encoder kernel internal error %#x" % kern.flags)

    def sync_host(self):
        """Host arrays as of the opcodes pulled so far (the generator stays live)."""
        v = self.v
        if self.synced_at == self.pulled:
            return
        if self.pulled == 0:
            slot = self.base
        else:
            kern = self._settle()
            self._check(kern)
            slot = kern.out
            self._advance_rng(kern)
        v._download_slot(slot)
        self.synced_at = self.pulled

    def _advance_rng(self, kern: _Kernel):
        """Move the process-global generators to where the reference's stand after the
        opcodes pulled so far (``kern`` = the run of exactly those opcodes, or one that ran
        dry before them)."""
        v = self.v
        v._py_mark = v._rng.set_states(kern.mt_np, kern.mt_py, self.gauss_next)
        v._np_mark = v._rng.numpy_mark()

    def finish(self):
        """The generator is closed, abandoned or garbage collected after ``pulled`` pulls:
        that state becomes the committed one and the global generators move on by the words
        the reference would have drawn."""
        if self.closed:
            return
        self.closed = True
        v = self.v
        if v._live is self:
            v._live = None
        v._pulls_closed += self.pulled
        if sys.is_finalizing():               # interpreter shutdown: CUDA may be gone
            return
        k = self.pulled
        v._predictor.pulled(k)
        if k == 0:
            v._release(self.kernel)
            v._drop_spec()
            return
        kern = self._settle()
        self._check(kern)
        if v._spec is not None and v._spec.base != kern.out:
            v._drop_spec()
        old = v._committed
        v._committed = kern.out
        if old != kern.out and old not in v._free_slots:
            v._free_slots.append(old)
        v._host_fresh = self.synced_at == k
        self._advance_rng(kern)

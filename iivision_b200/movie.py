"""Byte-stream side of the reference's transcoder/movie.py (SURVEY.md "next" row N2).

``Movie`` itself -- audio decoding, the frame grabber, the tick loop -- is out of scope;
what is here is what turns the encoder's opcode tuples into the player's byte stream:

* ``StreamMuxer.emit_stream(ops)``: the reference's generator (movie.py:122-161), one
  opcode at a time, same state (``stream_pos``, ``aux_memory_bank``, ``max_bytes_out``);
* ``emit_stream_device(...)``: the same bytes for a whole array of tick opcodes in one
  kernel (``iiv_emit_stream``) -- the form that composes with ``ops.encode_clips``;
* ``stream_schedule(...)``: the (frame, bank, budget) segments ``Movie.encode`` produces,
  including the first frame's 291 ticks (the 7-byte header shifts the first Ack).
"""

from typing import Iterable, Iterator, List, Tuple

import numpy as np
import torch

from . import ops
from . import opcodes
from ._lib import check, lib
from .machine import Machine
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
    """(frame, is_aux, budget) segments as Movie.encode + emit_stream produce them: a new
    encode_frame per encoded frame and, in DHGR, per bank flip -- after 291 tick opcodes
    in the first 2 KiB frame (7 header bytes), every 292 afterwards."""
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

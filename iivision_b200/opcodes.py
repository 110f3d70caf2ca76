"""Opcodes representing discrete operations of video player (reference
transcoder/opcodes.py).

Same classes and byte emission as the reference.  The player's opcode addresses come
from the cc65 debug file of the player build; like the reference this module reads
``player/iivision.dbg`` relative to the working directory at import time
(opcodes.py:173), or the file named by ``IIVISION_DBG``.  When neither exists the
addresses stay unset and ``load_symbols(path)`` can be called later; emitting an
address-bearing opcode before that raises ``ValueError`` (the reference raises at
import).  ``address_table()`` is what the batched device emitter
(``movie.emit_stream_device``) takes.
"""

import enum
import os
from typing import Iterator, Tuple

import numpy as np

from . import symbol_table
from . import video_mode
from .machine import Machine

TICKS = tuple(range(4, 68, 2))
PAGES = tuple(range(32, 64))


def _op_cmds():
    """Construct names of player opcodes."""
    op_cmds = ["HEADER", "TERMINATE", "NOP", "ACK"]
    for tick in TICKS:
        for page in PAGES:
            op_cmds.append("TICK_%d_PAGE_%d" % (tick, page))
    return op_cmds


OpcodeCommand = enum.Enum("OpcodeCommand", _op_cmds())


class Opcode:
    """One operation of the player.  A subclass names its command, the attributes that make
    up its payload (``FIELDS``, also the constructor's positional arguments) and how the
    payload is laid out in the stream (``_payload``); equality, repr and emission follow
    from those.  ``_START`` is the address of the player's implementation of the command."""
    COMMAND = None  # type: OpcodeCommand
    FIELDS = ()     # type: Tuple[str, ...]
    VECTORS = True  # whether the stream carries the address of this opcode before its data
    _START = None   # type: int

    def __init__(self, *values):
        if len(values) != len(self.FIELDS):
            raise TypeError("%s takes %d argument(s)" % (type(self).__name__, len(self.FIELDS)))
        for name, value in zip(self.FIELDS, values):
            setattr(self, name, value)

    def __repr__(self):
        return "Opcode(%s)" % self.COMMAND.name

    def __data_eq__(self, other):
        return all(getattr(self, f) == getattr(other, f) for f in self.FIELDS)

    def __eq__(self, other):
        return isinstance(other, self.__class__) and self.__data_eq__(other)

    __hash__ = None

    @staticmethod
    def emit_command(opcode: "Opcode") -> Iterator[int]:
        """The two address bytes (high, low) that vector the player to ``opcode``."""
        if not opcode.VECTORS:
            return
        if not opcode._START:
            raise ValueError(
                "Unable to find opcode address for %s in player debug symbols"
                % opcode.COMMAND)
        yield from divmod(opcode._START, 256)

    def _payload(self) -> Tuple[int, ...]:
        return ()

    def emit_data(self) -> Iterator[int]:
        yield from self._payload()

    def apply(self, state: Machine):
        pass


class Header(Opcode):
    """Video header: first bytes of the stream, not vectored to.  Padded with 0xff to the
    size of a tick opcode so that ACKs fall at even intervals; last byte = video mode."""
    COMMAND = OpcodeCommand.HEADER
    FIELDS = ("video_mode",)
    VECTORS = False

    def __init__(self, mode: video_mode.VideoMode):
        super().__init__(mode)

    def _payload(self):
        return (0xff,) * 6 + (self.video_mode.value,)


class Nop(Opcode):
    """Does nothing except vector to the next opcode."""
    COMMAND = OpcodeCommand.NOP


class Terminate(Opcode):
    """Ends playback."""
    COMMAND = OpcodeCommand.TERMINATE


class Ack(Opcode):
    """Player does its TCP stream + buffer management; carries the low byte of the
    $C054 / $C055 soft switch that steers the following stores to MAIN / AUX memory, and a
    pad byte that completes the TCP frame."""
    COMMAND = OpcodeCommand.ACK
    FIELDS = ("aux_active",)

    def __init__(self, aux_active: bool):
        super().__init__(aux_active)

    def _payload(self):
        return (0x55 if self.aux_active else 0x54, 0xff)


class BaseTick(Opcode):
    """"Fat" audio + video opcode, one class per (speaker duty cycle, HiRes page): stores
    the content byte at four offsets of that page."""
    FIELDS = ("content", "offsets")

    def __init__(self, content: int, offsets: Tuple):
        if len(offsets) != 4:
            raise ValueError("Wrong number of offsets: %d != 4" % len(offsets))
        super().__init__(content, offsets)

    def _payload(self):
        return (self.content,) + tuple(self.offsets)


TICK_OPCODES = {
    (_tick, _page): type("Tick%dPage%d" % (_tick, _page), (BaseTick,),
                         {"COMMAND": OpcodeCommand["TICK_%d_PAGE_%d" % (_tick, _page)]})
    for _tick in TICKS for _page in PAGES
}

_OPCODE_CLASSES = {
    OpcodeCommand.HEADER: Header,
    OpcodeCommand.TERMINATE: Terminate,
    OpcodeCommand.NOP: Nop,
    OpcodeCommand.ACK: Ack,
}
for (_tick, _page), _cls in TICK_OPCODES.items():
    _OPCODE_CLASSES[_cls.COMMAND] = _cls


def load_symbols(debugfile: str = "player/iivision.dbg") -> None:
    """Populate _START on every opcode class from the player's cc65 debug file."""
    by_name = {op.name.lower(): op for op in OpcodeCommand}
    found = {}
    for name, data in symbol_table.SymbolTable(debugfile).parse().items():
        if name.startswith("\"op_"):
            op = by_name.get(name[4:-1])
            if op is not None:
                found[op] = int(data["val"], 16)
    for op, cls in _OPCODE_CLASSES.items():
        if not found.get(op):
            raise ValueError(
                "Unable to find opcode address for %s in player debug symbols" % op)
    for op, start in found.items():
        _OPCODE_CLASSES[op]._START = start


def set_addresses(tick_addr, ack: int, terminate: int, header: int = 0, nop: int = 0) -> None:
    """Populate _START from an explicit table (tick_addr[32 ticks][32 pages])."""
    tick_addr = np.asarray(tick_addr).reshape(32, 32)
    for i, tick in enumerate(TICKS):
        for j, page in enumerate(PAGES):
            TICK_OPCODES[(tick, page)]._START = int(tick_addr[i, j])
    Ack._START, Terminate._START = int(ack), int(terminate)
    Header._START, Nop._START = int(header) or None, int(nop) or None


def address_table():
    """(uint16[32][32] tick addresses indexed [(tick-4)/2][page-32], ack, terminate)."""
    table = np.zeros((32, 32), dtype=np.uint16)
    for i, tick in enumerate(TICKS):
        for j, page in enumerate(PAGES):
            start = TICK_OPCODES[(tick, page)]._START
            if not start:
                raise ValueError("opcode addresses not loaded (opcodes.load_symbols)")
            table[i, j] = start
    if not Ack._START or not Terminate._START:
        raise ValueError("opcode addresses not loaded (opcodes.load_symbols)")
    return table, int(Ack._START), int(Terminate._START)


_dbg = os.environ.get("IIVISION_DBG", "player/iivision.dbg")
if os.path.exists(_dbg):
    load_symbols(_dbg)

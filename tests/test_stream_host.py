"""CPU: host side of the byte-stream row (N2) -- StreamMuxer / opcodes / symbol table --
against bytes produced by the reference's Movie.emit_stream (tests/golden/
byte_stream.npz, generator oracle/make_golden.py)."""

import io
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["short", "frames", "exact", "capped"]


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "byte_stream.npz"))


@pytest.fixture(scope="module")
def mods(g):
    from iivision_b200 import _build
    _build.build()
    from iivision_b200 import movie, opcodes, symbol_table, video_mode
    opcodes.set_addresses(g["tick_addr"], int(g["ack_addr"]), int(g["terminate_addr"]))
    import types
    return types.SimpleNamespace(movie=movie, opcodes=opcodes, symbol_table=symbol_table,
                                 video_mode=video_mode)


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
@pytest.mark.parametrize("case", CASES)
def test_stream_muxer_matches_reference_bytes(mods, g, mode, case):
    key = "%s_%s" % (mode.lower(), case)
    rec, ticks = g[key + "_records"], g[key + "_ticks"]
    vm = getattr(mods.video_mode.VideoMode, mode)
    mux = mods.movie.StreamMuxer(vm, max_bytes_out=int(g[key + "_max"]) or None)
    ops = [mods.opcodes.Header(mode=vm)] + [
        mods.opcodes.TICK_OPCODES[(int(ticks[k]), int(rec[k, 0]))](
            int(rec[k, 1]), tuple(int(x) for x in rec[k, 2:6])) for k in range(len(rec))]
    data = bytes(mux.emit_stream(ops))
    assert data == g[key + "_bytes"].tobytes()
    assert len(data) % 2048 == 0


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
@pytest.mark.parametrize("case", CASES)
def test_stream_length_arithmetic(mods, g, mode, case):
    from iivision_b200._lib import lib
    key = "%s_%s" % (mode.lower(), case)
    n = len(g[key + "_records"])
    n_emit = lib.iiv_stream_ticks_within(n, int(g[key + "_max"]))
    assert lib.iiv_stream_length(n_emit, 1) == len(g[key + "_bytes"])
    if case != "capped":
        assert n_emit == n


def test_opcode_classes_and_symbols(mods):
    oc = mods.opcodes
    assert len(oc.OpcodeCommand) == 4 + 32 * 32
    a = oc.TICK_OPCODES[(4, 32)](1, (2, 3, 4, 5))
    assert a == oc.TICK_OPCODES[(4, 32)](1, (2, 3, 4, 5)) and a != oc.TICK_OPCODES[(4, 33)](1, (2, 3, 4, 5))
    assert a != oc.TICK_OPCODES[(4, 32)](1, (2, 3, 4, 6)) and oc.Nop() == oc.Nop()
    assert oc.Ack(True) != oc.Ack(False) and repr(oc.Terminate()) == "Opcode(TERMINATE)"
    with pytest.raises(ValueError):
        oc.TICK_OPCODES[(4, 32)](1, (2, 3, 4))
    # cc65 debug-file parsing (the format symbol_table_test.py feeds the reference)
    dbg = io.StringIO(
        'version\tmajor=2,minor=0\n'
        'sym\tid=0,name="op_ack",addr_size=absolute,size=1,scope=0,def=1,val=0xBA72,seg=0,type=lab\n'
        'sym\tid=1,name="other",addr_size=absolute,scope=0,def=2,val=0x1234,seg=0,type=lab\n')
    syms = mods.symbol_table.SymbolTable().parse(dbg)
    assert syms['"op_ack"']["val"] == "0xBA72" and len(syms) == 2
    table, ack, term = oc.address_table()
    assert table.shape == (32, 32) and ack == oc.Ack._START


def test_stream_schedule():
    from iivision_b200 import movie
    segs = movie.stream_schedule("DHGR", 2)
    assert segs[0] == (0, 0, 291) and segs[1] == (0, 1, 292)
    assert sum(s[2] for s in segs) == 1960
    # the bank changes exactly when the running count hits 291 + 292 k; a new frame
    # starts a new segment on the bank in force
    count, aux = 0, 0
    for frame, is_aux, budget in segs:
        assert is_aux == aux
        count += budget
        if (count - 291) % 292 == 0 and count >= 291:
            aux ^= 1
    assert [s[0] for s in segs] == sorted(s[0] for s in segs)
    assert movie.stream_schedule("HGR", 3) == [(0, 0, 980), (1, 0, 980), (2, 0, 980)]

"""Parity of the CUDA table generator against the CPU oracle (bit-exact)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PALETTES = {"NTSC": 5, "IIGS": 0}


@pytest.fixture(scope="module")
def ops():
    from iivision_b200 import ops
    return ops


@pytest.mark.parametrize("pid", [5, 0])
def test_lut_matches_oracle(ops, oracle_luts, pid):
    from oracle import cie2000, palettes
    got = ops.lut_cie2000(palettes.RGB[pid])
    assert np.array_equal(got, oracle_luts[pid])
    # untruncated values agree to far better than the smallest margin to an
    # integer (1.5e-5 at black<->white)
    f_dev = ops.lut_cie2000_float(palettes.RGB[pid])
    f_cpu = cie2000.diff_matrix_float(palettes.RGB[pid])
    assert np.max(np.abs(f_dev - f_cpu)) < 1e-9
    assert got[0, 15] == 99


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_dots_and_pixel_strings(ops, mode):
    from oracle import tables
    assert np.array_equal(ops.all_dots(mode).cpu().numpy().view(np.uint32),
                          tables.all_dots(mode))
    assert np.array_equal(ops.all_pixel_strings(mode).cpu().numpy(),
                          tables.all_pixel_strings(mode))


@pytest.mark.parametrize("mode,pid", [("DHGR", 5), ("DHGR", 0), ("HGR", 5), ("HGR", 0)])
def test_full_table_bit_exact(ops, oracle_luts, mode, pid):
    """Every entry of the reference-layout (lower-triangular) table."""
    from oracle import tables
    want, n = tables.build_table(mode, oracle_luts[pid], triangular=True)
    got = ops.table_generate(mode, oracle_luts[pid], layout=ops.LAYOUT_TRIANGULAR)
    got = got.cpu().numpy()
    mism = int(np.count_nonzero(got != want))
    assert mism == 0, "%d of %d entries differ" % (mism, want.size)
    bits = tables.MASKED_BITS[mode]
    assert n == tables.NUM_OFFSETS[mode] * (1 << bits) * ((1 << bits) - 1) // 2


@pytest.mark.parametrize("mode", ["DHGR", "HGR"])
@pytest.mark.parametrize("layout", [0, 1])
def test_tree_kernel_equals_chain_kernel(ops, oracle_luts, mode, layout):
    """The shared-suffix tree generator against the one-chain-per-entry kernel,
    with an adversarial LUT (zeros off the diagonal, values up to 255)."""
    import torch
    from iivision_b200._lib import ALGO_CHAIN, ALGO_SPLIT, ALGO_TREE
    rng = np.random.default_rng(11)
    lut = rng.integers(0, 256, size=(16, 16)).astype(np.int32)
    lut = np.minimum(lut, lut.T)
    lut[rng.random((16, 16)) < 0.15] = 0
    lut = np.minimum(lut, lut.T)
    np.fill_diagonal(lut, 0)
    for table_lut in (oracle_luts[5], lut):
        a = ops.table_generate(mode, table_lut, layout=layout, algo=ALGO_CHAIN)
        b = ops.table_generate(mode, table_lut, layout=layout, algo=ALGO_TREE)
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))
        del b
        c = ops.table_generate(mode, table_lut, layout=layout, algo=ALGO_SPLIT)
        assert torch.equal(a.view(torch.int16), c.view(torch.int16))


def test_split_kernel_extreme_luts(ops):
    """The split generator's packed 16-bit arithmetic at its limits: every substitution
    costs 255 (entries up to n * 255, the INF marker of an absent swap must never win),
    and a LUT of zeros (every entry 0 or a swap's 1)."""
    import torch
    from iivision_b200._lib import ALGO_CHAIN, ALGO_SPLIT
    hi = np.full((16, 16), 255, dtype=np.int32)
    np.fill_diagonal(hi, 0)
    for lut in (hi, np.full((16, 16), 255, dtype=np.int32), np.zeros((16, 16), dtype=np.int32)):
        for mode in ("HGR", "DHGR"):
            a = ops.table_generate(mode, lut, layout=1, algo=ALGO_CHAIN)
            c = ops.table_generate(mode, lut, layout=1, algo=ALGO_SPLIT)
            assert torch.equal(a.view(torch.int16), c.view(torch.int16))
            del a, c


def test_concurrent_generates_on_two_streams(ops, oracle_luts):
    """Two generate calls with different LUTs in flight at once (the split generator's
    scratch tables are per call, not globals)."""
    import torch
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    want = {pid: ops.table_generate("DHGR", oracle_luts[pid]) for pid in (5, 0)}
    outs = {pid: torch.empty_like(want[pid]) for pid in (5, 0)}
    torch.cuda.synchronize()
    for _ in range(4):
        with torch.cuda.stream(s1):
            ops.table_generate("DHGR", oracle_luts[5], out=outs[5])
        with torch.cuda.stream(s2):
            ops.table_generate("DHGR", oracle_luts[0], out=outs[0])
    torch.cuda.synchronize()
    for pid in (5, 0):
        assert torch.equal(outs[pid].view(torch.int16), want[pid].view(torch.int16))


def test_row_ranges_unaligned(ops, oracle_luts):
    """Row ranges that do not fall on the generator's 8-row tiles."""
    import torch
    lut = oracle_luts[0]
    for layout in (0, 1):
        full = ops.table_generate("HGR", lut, layout=layout)
        out = torch.zeros_like(full)
        for a, b in ((0, 3), (3, 8), (8, 9), (9, 1023), (1023, 1025), (1025, 16381),
                     (16381, 16384)):
            ops.table_generate("HGR", lut, layout=layout, row_begin=a, row_end=b, out=out)
        assert torch.equal(out.view(torch.int16), full.view(torch.int16))


@pytest.mark.parametrize("mode", ["DHGR", "HGR"])
def test_symmetrise_and_symmetric_layout(ops, oracle_luts, oracle_tables, mode):
    lut = oracle_luts[5]
    tri = ops.table_generate(mode, lut, layout=ops.LAYOUT_TRIANGULAR)
    ops.table_symmetrise(mode, tri)
    sym = ops.table_generate(mode, lut, layout=ops.LAYOUT_SYMMETRIC)
    import torch
    assert torch.equal(tri.view(torch.int16), sym.view(torch.int16))
    assert np.array_equal(sym.cpu().numpy(), oracle_tables(mode, 5))


def test_row_blocks_compose(ops, oracle_luts):
    """Shards by row block (the multi-GPU partition) tile the full table."""
    import torch
    lut = oracle_luts[5]
    full = ops.table_generate("DHGR", lut, layout=ops.LAYOUT_SYMMETRIC)
    out = torch.zeros_like(full)
    edges = [0, 1, 1000, 1024, 4097, 8192]
    for a, b in zip(edges[:-1], edges[1:]):
        ops.table_generate("DHGR", lut, layout=ops.LAYOUT_SYMMETRIC, row_begin=a,
                           row_end=b, out=out)
    assert torch.equal(out.view(torch.int16), full.view(torch.int16))
    # empty range is a no-op
    ops.table_generate("DHGR", lut, row_begin=5, row_end=5, out=out)


def test_known_answers(ops, device_tables):
    """Facts recorded in SURVEY.md 8(c) for the NTSC DHGR table."""
    t = device_tables("DHGR", 5).cpu().numpy()
    assert t.max() == 1010
    assert t[0][(0 << 13) + 0x1FFF] == 990
    for o in range(4):
        zeros = np.flatnonzero(t[o] == 0)
        off_diag = [z for z in zeros if (z >> 13) != (z & 0x1FFF)]
        assert sorted(off_diag) == sorted(
            [(0x0AAA << 13) + 0x1555, (0x1555 << 13) + 0x0AAA])


def test_reference_invariants_iigs(ops, device_tables):
    """make_data_tables_test.py:18-53: symmetric, >= 0, zero only on the diagonal
    (holds for IIGS; NTSC has GREY1 == GREY2, see test_known_answers)."""
    t = device_tables("DHGR", 0).cpu().numpy()
    for o in range(4):
        sq = t[o].reshape(8192, 8192)
        assert np.array_equal(sq, sq.T)
        zi, zj = np.nonzero(sq == 0)
        assert np.array_equal(zi, zj) and len(zi) == 8192


def test_string_distance(ops, oracle_luts):
    from oracle import tables
    rng = np.random.default_rng(7)
    a = rng.integers(0, 16, size=(500, 18), dtype=np.uint8)
    b = rng.integers(0, 16, size=(500, 18), dtype=np.uint8)
    b[:100] = a[:100]
    b[100:200, 3:5] = a[100:200, 3:5][:, ::-1]      # plant transpositions
    got = ops.string_distance(oracle_luts[0], a, b)
    want = [tables.chain_distance(a[k], b[k], oracle_luts[0]) for k in range(500)]
    assert got.tolist() == want


def test_bad_arguments(ops, oracle_luts):
    from iivision_b200._lib import IIVError
    import torch
    lut = oracle_luts[5].copy()
    out = torch.empty(ops.table_shape("DHGR"), dtype=torch.uint16, device="cuda")
    with pytest.raises(IIVError):
        ops.table_generate("DHGR", lut, row_begin=9000, row_end=9001, out=out)
    lut[3, 4] = 300
    with pytest.raises(IIVError):
        ops.table_generate("DHGR", lut, out=out)

"""GPU: the CUDA path against fixtures produced by the UNMODIFIED reference
(tests/golden; generator oracle/make_golden.py) -- no oracle in the loop except
for the edit-distance table the reference was fed."""

import hashlib
import json
import os

import numpy as np
import pytest

from encoder_util import run_device

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ops():
    from iivision_b200 import ops
    return ops


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_pixel_strings_reference_hashes(ops, mode):
    g = json.load(open(os.path.join(GOLDEN, "pixel_strings.json")))[mode]
    assert _sha(ops.all_dots(mode).cpu().numpy().view(np.uint32)) == g["dots_sha256"]
    assert _sha(ops.all_pixel_strings(mode).cpu().numpy()) == g["pixels_sha256"]


@pytest.mark.parametrize("mode,pid", [("HGR", 5), ("HGR", 0), ("DHGR", 5), ("DHGR", 0)])
def test_table_hashes(ops, oracle_luts, mode, pid):
    g = json.load(open(os.path.join(GOLDEN, "luts.json")))
    from oracle import palettes
    lut = ops.lut_cie2000(palettes.RGB[pid])
    assert lut.tolist() == g["lut"][str(pid)]
    tab = ops.table_generate(mode, lut, layout=ops.LAYOUT_TRIANGULAR)
    assert _sha(tab.cpu().numpy()) == g["table_sha256"]["%s_%d" % (mode, pid)]


@pytest.mark.parametrize("mode", ["HGR", "DHGR"])
def test_scorer_primitives(ops, device_tables, mode):
    import torch
    g = np.load(os.path.join(GOLDEN, "scorer_%s.npz" % mode.lower()))
    table = device_tables(mode)
    fr = g["frames"]
    main = _dev(fr[:, 0])
    aux = _dev(fr[:, 1]) if mode == "DHGR" else None
    packed = ops.pack(mode, main, aux)
    assert np.array_equal(packed[0].cpu().numpy().view(np.uint64), g["src_packed"])
    assert np.array_equal(packed[1].cpu().numpy().view(np.uint64), g["tgt_packed"])
    words = _dev(g["words"].view(np.int64))
    for o in range(g["mask_shift"].shape[0]):
        got = ops.mask_and_shift(mode, o, words).cpu().numpy().view(np.uint64)
        assert np.array_equal(got, g["mask_shift"][o])
        for k, v in enumerate(g["values"]):
            got = ops.masked_update(mode, o, words, int(v)).cpu().numpy().view(np.uint64)
            assert np.array_equal(got, g["masked_update"][o, k])
    src, tgt = packed[0].contiguous(), packed[1].contiguous()
    for tag in (("main", "aux") if mode == "DHGR" else ("main",)):
        is_aux = tag == "aux"
        dw = ops.diff_weights(mode, is_aux, src, tgt, table)
        assert np.array_equal(dw.cpu().numpy(), g["diff_weights_" + tag])
        for (page, content), want in zip(g["delta_cases_" + tag], g["delta_" + tag]):
            got = ops.compute_delta_page(mode, is_aux, tgt, int(page), int(content),
                                         dw[int(page)].contiguous(), table)
            assert np.array_equal(got.cpu().numpy(), want)
            row = tgt[int(page)].contiguous()
            nd = ops.diff_weights_page(mode, is_aux, row, row, table, int(content))
            assert np.array_equal(nd.cpu().numpy() - dw[int(page)].cpu().numpy(), want)
        rows = ops.delta_rows(mode, is_aux, tgt, table).cpu().numpy().astype(np.int32)
        for (page, content), want in zip(g["delta_cases_" + tag], g["delta_" + tag]):
            assert np.array_equal(rows[page, content] - g["diff_weights_" + tag][page], want)
        for bo, page, off, content, want in g["pair_difference_" + tag]:
            got = ops.byte_pair_difference(
                mode, int(bo), tgt[page, off // 2].reshape(1).contiguous(),
                torch.tensor([content], dtype=torch.uint8, device="cuda"), table)
            assert int(got.cpu().numpy().view(np.uint16)[0]) == want
    p = src.clone()
    m = main[0].clone()
    a = aux[0].clone() if mode == "DHGR" else None
    ops.apply_stores(mode, p, m, a, g["apply_stores"])
    assert np.array_equal(p.cpu().numpy().view(np.uint64), g["apply_packed"])
    assert np.array_equal(m.cpu().numpy(), g["apply_main"])
    if mode == "DHGR":
        assert np.array_equal(a.cpu().numpy(), g["apply_aux"])


@pytest.mark.parametrize("name", ["dhgr_full", "dhgr_sparse", "hgr_full", "hgr_sparse",
                                  "hgr_long_generator", "hgr_exhaust", "dhgr_long_generator"])
def test_opcode_streams(ops, device_tables, name):
    import torch
    g = np.load(os.path.join(GOLDEN, "stream_%s.npz" % name))
    mode = str(g["mode"])
    seed = int(g["rng_seed"])
    got, info, states = run_device(ops, mode, device_tables(mode), g["frames"][None],
                                   g["segments"], [seed])
    assert np.array_equal(got[0][:, :6], g["opcodes"])
    assert np.array_equal(got[0][:, 6], g["real"])
    # video.py:90 "Similarity": mean of update_priority before each segment
    assert np.array_equal(info[0][:, 1] / 8192.0, g["similarity"])
    packed = ops.state_field(states, ops.F_PACKED, torch.int64, (32, 128))[0]
    assert np.array_equal(packed.cpu().numpy().view(np.uint64), g["packed"])
    assert np.array_equal(
        ops.state_field(states, ops.F_MAIN, torch.uint8, (32, 256))[0].cpu().numpy(), g["main"])
    assert np.array_equal(
        ops.state_field(states, ops.F_PRIO_MAIN, torch.int32, (32, 256))[0].cpu().numpy(),
        g["priority_main"])
    if mode == "DHGR":
        assert np.array_equal(
            ops.state_field(states, ops.F_AUX, torch.uint8, (32, 256))[0].cpu().numpy(), g["aux"])
        assert np.array_equal(
            ops.state_field(states, ops.F_PRIO_AUX, torch.int32, (32, 256))[0].cpu().numpy(),
            g["priority_aux"])
    # RNG streams: next words equal what the reference's global generators gave
    mt_py = ops.state_field(states, ops.F_MT_PY, torch.int32, (640,))[0].clone()
    nxt = ops.mt_draw(mt_py, 4).cpu().numpy().view(np.uint32)
    assert nxt.tolist() == g["next_python_words"].tolist()
    mt_np = ops.state_field(states, ops.F_MT_NP, torch.int32, (640,))[0].clone()
    nxt = ops.mt_draw(mt_np, 4).cpu().numpy().view(np.uint32) & 0xFF
    assert nxt.tolist() == g["next_numpy_bytes"].tolist()

"""BASELINE.json configs[3]: one 6000-frame synthetic 560x192 DHGR clip on one B200
(a single clip is sequential encoder state: one thread block)."""
import random, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from iivision_b200 import ops, synth, palette
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
lut = ops.lut_cie2000(palette.NTSCPalette.rgb_by_value())
table = ops.table_generate("DHGR", lut, layout=ops.LAYOUT_SYMMETRIC)
t0 = time.time(); frames = synth.synthetic_frames("DHGR", n_frames, 1.0, seed=5); print("frames generated in %.1f s" % (time.time() - t0))
plan = ops.SegmentPlan(synth.movie_schedule("DHGR", n_frames))
tmem = torch.from_numpy(frames[None]).cuda()
tpacked = ops.pack("DHGR", tmem[0, :, 0].contiguous(), tmem[0, :, 1].contiguous()).view(1, n_frames, 32, 128)
def fresh():
    st = ops.new_clip_states(1)
    pad = np.zeros(640, np.uint32)
    pad[:625] = ops.mt_from_python(random.Random(0).getstate())
    ops.state_field(st, ops.F_MT_PY, torch.int32, (640,)).copy_(torch.from_numpy(pad.view(np.int32)).cuda().expand(1, 640))
    pad[:625] = ops.mt_from_numpy(np.random.RandomState(0).get_state())
    ops.state_field(st, ops.F_MT_NP, torch.int32, (640,)).copy_(torch.from_numpy(pad.view(np.int32)).cuda().expand(1, 640))
    return st
opc = torch.empty((1, plan.total, 8), dtype=torch.uint8, device="cuda")
info = torch.zeros((1, len(plan), 8), dtype=torch.int64, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
st = fresh(); torch.cuda.synchronize(); ev[0].record()
ops.encode_clips("DHGR", st, tmem, tpacked, plan, table, opcodes=opc, seg_info=info)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1])
inf = info.cpu().numpy()[0]
print("%d frames, %d opcodes (%d real) in %.1f ms: %.1f frames/s, %.3f us/opcode; flags %s" % (
    n_frames, plan.total, int(inf[:, 0].sum()), ms, n_frames / (ms * 1e-3), ms * 1e3 / plan.total,
    ops.state_field(st, ops.F_FLAGS, torch.int32, (8,))[0].cpu().numpy().tolist()))
# prefix check against a short run and the oracle
short = 3
plan_s = ops.SegmentPlan(synth.movie_schedule("DHGR", short))
opc_s = torch.empty((1, plan_s.total, 8), dtype=torch.uint8, device="cuda")
info_s = torch.zeros((1, len(plan_s), 8), dtype=torch.int64, device="cuda")
ops.encode_clips("DHGR", fresh(), tmem[:, :short].contiguous(), tpacked[:, :short].contiguous(), plan_s, table, opcodes=opc_s, seg_info=info_s)
torch.cuda.synchronize()
print("prefix of the long run == 3-frame run:", bool(torch.equal(opc[0, :plan_s.total], opc_s[0])))
sys.path.insert(0, 'tests')
from encoder_util import run_oracle
want, *_ = run_oracle("DHGR", table.cpu().numpy(), frames[:short], synth.movie_schedule("DHGR", short), 0)
print("3-frame run == oracle:", bool(np.array_equal(opc_s.cpu().numpy()[0][:, :6].astype(np.int64), want)))

import sys, torch, time
sys.path.insert(0, '.')
from iivision_b200 import ops, synth, palette, _lib
lut = ops.lut_cie2000(palette.NTSCPalette.rgb_by_value())
table = ops.table_generate("DHGR", lut, layout=ops.LAYOUT_SYMMETRIC)
nb = 1024
fr = synth.synthetic_frames("DHGR", nb + 1, 1.0, seed=1)
d = torch.from_numpy(fr).cuda()
src = ops.pack("DHGR", d[:nb, 0].contiguous(), d[:nb, 1].contiguous())
tgt = d[1:].contiguous()
prio = torch.zeros((nb, 2, 32, 256), dtype=torch.int32, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
def t_score():
    for _ in range(3): ops.score_frames("DHGR", src, tgt, table, priority=prio)
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(20): ops.score_frames("DHGR", src, tgt, table, priority=prio)
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / 20
import numpy as np
clips = np.stack([synth.synthetic_frames("DHGR", 4, 1.0, seed=100 + c) for c in range(148)])
plan = ops.SegmentPlan(synth.movie_schedule("DHGR", 4))
tmem = torch.from_numpy(clips).cuda(); flat = tmem.view(-1, 2, 32, 256)
tpacked = ops.pack("DHGR", flat[:, 0].contiguous(), flat[:, 1].contiguous()).view(148, 4, 32, 128)
opc = torch.empty((148, plan.total, 8), dtype=torch.uint8, device="cuda")
info = torch.zeros((148, len(plan), 8), dtype=torch.int64, device="cuda")
def t_enc():
    ts = []
    for r in range(7):
        st = ops.seed_clip_states(ops.new_clip_states(148), range(148))
        torch.cuda.synchronize(); ev[0].record()
        ops.encode_clips("DHGR", st, tmem, tpacked, plan, table, opcodes=opc, seg_info=info)
        ev[1].record(); torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]))
    return sorted(ts)[3]
print("default granularity", _lib.lib.iiv_get_l2_fetch_granularity())
for g in (0, 32, 64, 128, 32):
    if g: _lib.check(_lib.lib.iiv_set_l2_fetch_granularity(g))
    print("gran", g or "default", _lib.lib.iiv_get_l2_fetch_granularity(), "score_frames ms %.4f" % t_score(), "encode 148x4 ms %.3f" % t_enc(), flush=True)

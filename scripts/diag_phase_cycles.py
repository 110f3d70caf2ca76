import numpy as np, torch, random, sys
sys.path.insert(0,'.')
from iivision_b200 import ops, synth, palette
lut = ops.lut_cie2000(palette.NTSCPalette.rgb_by_value())
table = ops.table_generate("DHGR", lut, layout=ops.LAYOUT_SYMMETRIC)
n_frames=int(sys.argv[1]) if len(sys.argv) > 1 else 4
clips = synth.synthetic_frames("DHGR", n_frames, 1.0, seed=100)[None]
segs = synth.movie_schedule("DHGR", n_frames)
tmem = torch.from_numpy(np.ascontiguousarray(clips)).cuda()
flat = tmem.view(-1, 2, 32, 256)
tpacked = ops.pack("DHGR", flat[:, 0].contiguous(), flat[:, 1].contiguous()).view(1, n_frames, 32, 128)
for rep in range(3):
    st = ops.new_clip_states(1)
    pad = np.zeros(640, np.uint32); pad[:625] = ops.mt_from_python(random.Random(0).getstate())
    ops.state_field(st, ops.F_MT_PY, torch.int32, (640,)).copy_(torch.from_numpy(pad.view(np.int32)).cuda().expand(1, 640))
    pad[:625] = ops.mt_from_numpy(np.random.RandomState(0).get_state())
    ops.state_field(st, ops.F_MT_NP, torch.int32, (640,)).copy_(torch.from_numpy(pad.view(np.int32)).cuda().expand(1, 640))
    _, info = ops.encode_clips("DHGR", st, tmem, tpacked, segs, table)
    torch.cuda.synchronize()
info = info.cpu().numpy()[0]
print("seg: emitted n_heap | total_A | loop | wait_rows wait_mt wait_applier")
for k in range(info.shape[0]):
    r = info[k]
    print(k, r[0], r[2], "|", r[4], "|", r[5], "|", r[6] & 0xffffffff, r[6] >> 32, "twist", r[7] & 0xffffffff, "fill", r[7] >> 32)

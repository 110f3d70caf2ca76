import sys, random, numpy as np, torch
sys.path.insert(0,'.')
from iivision_b200 import ops, synth, palette
lut = ops.lut_cie2000(palette.NTSCPalette.rgb_by_value())
table = ops.table_generate("DHGR", lut, layout=ops.LAYOUT_SYMMETRIC)
n_clips, n_frames = 148, 4
clips = np.stack([synth.synthetic_frames("DHGR", n_frames, 1.0, seed=100 + c) for c in range(4)])
clips = np.concatenate([clips] * (n_clips // 4 + 1))[:n_clips]
segs = synth.movie_schedule("DHGR", n_frames)
tmem = torch.from_numpy(np.ascontiguousarray(clips)).cuda()
flat = tmem.view(-1, 2, 32, 256)
tpacked = ops.pack("DHGR", flat[:, 0].contiguous(), flat[:, 1].contiguous()).view(n_clips, n_frames, 32, 128)
mt_py = ops.mt_from_python(random.Random(0).getstate()); mt_np = ops.mt_from_numpy(np.random.RandomState(0).get_state())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for rep in range(12):
    st = ops.new_clip_states(n_clips)
    pad = np.zeros(640, np.uint32); pad[:625] = mt_py
    ops.state_field(st, ops.F_MT_PY, torch.int32, (640,)).copy_(torch.from_numpy(pad.view(np.int32)).cuda().expand(n_clips, 640))
    pad[:625] = mt_np
    ops.state_field(st, ops.F_MT_NP, torch.int32, (640,)).copy_(torch.from_numpy(pad.view(np.int32)).cuda().expand(n_clips, 640))
    torch.cuda.synchronize(); ev[0].record()
    _, info = ops.encode_clips("DHGR", st, tmem, tpacked, segs, table)
    ev[1].record(); torch.cuda.synchronize()
    inf = info.cpu().numpy()            # [clip][seg][8]
    tot = inf[:, :, 4].sum(1) + inf[:, :, 5].sum(1)
    k = int(tot.argmax())
    print("run %2d: %.2f ms | per-clip cycles min %.2fM med %.2fM max %.2fM (clip %d: A %.2fM loop %.2fM wait_rows %.2fM wait_mt %.2fM) | worst seg loop %d of clip %d" % (
        rep, ev[0].elapsed_time(ev[1]), tot.min()/1e6, np.median(tot)/1e6, tot.max()/1e6, k,
        inf[k,:,4].sum()/1e6, inf[k,:,5].sum()/1e6, inf[k,:,6].sum()/1e6, inf[k,:,7].sum()/1e6,
        inf[:,:,5].max(), int(inf[:,:,5].max(1).argmax())))

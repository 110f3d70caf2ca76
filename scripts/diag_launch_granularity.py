"""Does an encoder segment cost more when it is a launch of its own?  The same clip and
schedule once as ONE launch of all segments and once as one launch per segment (what the
generator facade does): device cycles of scoring + heapify and of the opcode loop, summed."""
import os, random, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iivision_b200 import ops, palette, synth

lut = ops.lut_cie2000(palette.NTSCPalette.rgb_by_value())
table = ops.table_generate("DHGR", lut, layout=ops.LAYOUT_SYMMETRIC)
n_frames = 6
clips = synth.synthetic_frames("DHGR", n_frames, 1.0, seed=100)[None]
segs = synth.movie_schedule("DHGR", n_frames)
tmem = torch.from_numpy(np.ascontiguousarray(clips)).cuda()
flat = tmem.view(-1, 2, 32, 256)
tpacked = ops.pack("DHGR", flat[:, 0].contiguous(), flat[:, 1].contiguous()).view(1, n_frames, 32, 128)


def fresh_state():
    st = ops.new_clip_states(1)
    pad = np.zeros(640, np.uint32)
    pad[:625] = ops.mt_from_python(random.Random(0).getstate())
    ops.state_field(st, ops.F_MT_PY, torch.int32, (640,)).copy_(
        torch.from_numpy(pad.view(np.int32)).cuda().expand(1, 640))
    pad[:625] = ops.mt_from_numpy(np.random.RandomState(0).get_state())
    ops.state_field(st, ops.F_MT_NP, torch.int32, (640,)).copy_(
        torch.from_numpy(pad.view(np.int32)).cuda().expand(1, 640))
    return st


for rep in range(2):
    st = fresh_state()
    out, info = ops.encode_clips("DHGR", st, tmem, tpacked, segs, table)
    torch.cuda.synchronize()
    one = info.cpu().numpy()[0]
    st = fresh_state()
    parts, outs = [], []
    for s in segs:
        o, i = ops.encode_clips("DHGR", st, tmem, tpacked, [s], table)
        torch.cuda.synchronize()
        parts.append(i.cpu().numpy()[0][0])
        outs.append(o.cpu().numpy()[0])
    parts = np.array(parts)
    same = np.array_equal(np.concatenate(outs), out.cpu().numpy()[0])
    n_op = int(one[:, 0].sum())
    print("run %d: %d segments, %d opcodes, streams equal: %s" % (rep, len(segs), n_op, same))
    print("  one launch        : score+heapify %8d cycles/segment, loop %6.0f cycles/opcode"
          % (one[:, 4].mean(), one[:, 5].sum() / n_op))
    print("  launch per segment: score+heapify %8d cycles/segment, loop %6.0f cycles/opcode"
          % (parts[:, 4].mean(), parts[:, 5].sum() / n_op))

"""Stress the encoder for nondeterminism: run the same launch repeatedly and
report where runs differ from the first one / from the oracle."""
import sys, random
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from iivision_b200 import ops, synth, palette
from encoder_util import run_device, run_oracle
mode = "DHGR"
lut = ops.lut_cie2000(palette.NTSCPalette.rgb_by_value())
table = ops.table_generate(mode, lut, layout=ops.LAYOUT_SYMMETRIC)
clips = np.stack([synth.synthetic_frames(mode, 2, f, seed=s)
                  for f, s in ((1.0, 10), (0.3, 11), (0.05, 12), (1.0, 13), (0.0, 14))])
segs = synth.movie_schedule(mode, 2, opcodes_per_frame=400, flip_every=150)
seeds = [100, 101, 102, 103, 104]
ref = None
host_table = table.cpu().numpy()
want = [run_oracle(mode, host_table, clips[k], segs, seeds[k])[0] for k in range(5)]
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    got, info, _ = run_device(ops, mode, table, clips, segs, seeds)
    for k in range(5):
        g = got[k][:, :6].astype(np.int64)
        if not np.array_equal(g, want[k]):
            bad += 1
            d = np.flatnonzero((g != want[k]).any(axis=1))
            # which segment
            edges = np.cumsum([0] + [s[2] for s in segs])
            seg = int(np.searchsorted(edges, d[0], side='right') - 1)
            print("iter", it, "clip", k, "first diff at opcode", int(d[0]), "of", len(g), "segment", seg, segs[seg],
                  "pos in seg", int(d[0] - edges[seg]), "got", g[d[0]].tolist(), "want", want[k][d[0]].tolist(),
                  "n_diff", len(d), "info", info[k][seg][:4].tolist())
print("bad clip-runs:", bad)

"""HGR scoring prologue, table gathers vs factor tables (the bench quotes DHGR)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iivision_b200 import ops, palette, synth  # noqa: E402

lut = ops.lut_cie2000(palette.NTSCPalette.rgb_by_value())
table = ops.table_generate("HGR", lut)
factors = ops.score_factors("HGR", lut)
nb = 1024
for fraction in (1.0, 0.05):
    fr = synth.synthetic_frames("HGR", nb + 1, fraction, seed=1)
    d = torch.from_numpy(fr).cuda()
    src = ops.pack("HGR", d[:nb, 0].contiguous(), None)
    tgt = d[1:].contiguous()
    out = {}
    for name, kw in (("table", {"table": table}), ("factors", {"factors": factors})):
        prio = torch.zeros((nb, 1, 32, 256), dtype=torch.int32, device="cuda")
        for _ in range(3):
            res = ops.score_frames("HGR", src, tgt, priority=prio, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.score_frames("HGR", src, tgt, priority=prio, **kw)
        e1.record()
        torch.cuda.synchronize()
        out[name] = (e0.elapsed_time(e1) / 20, res[1])
    same = torch.equal(out["table"][1], out["factors"][1])
    print("HGR fraction %.2f: table gathers %.3f ms (%.2f M frames/s), factors %.3f ms (%.2f M frames/s), "
          "same diff weights: %s" % (fraction, out["table"][0], nb / out["table"][0] / 1e3,
                                     out["factors"][0], nb / out["factors"][0] / 1e3, same))

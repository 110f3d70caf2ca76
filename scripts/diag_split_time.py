"""Times the table generator's step (all launches of one iiv_table_generate call) for the
HGR and DHGR tables; used by scripts/gpu_split_variants.sh."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iivision_b200 import ops  # noqa: E402

lut = np.random.default_rng(0).integers(0, 100, (16, 16)).astype(np.int32)
lut = np.minimum(lut, lut.T)
np.fill_diagonal(lut, 0)
res = []
for mode in ("HGR", "DHGR"):
    out = torch.empty(ops.table_shape(mode), dtype=torch.uint16, device="cuda")
    ref = ops.table_generate(mode, lut, algo=2)
    for layout in (1, 0):
        for _ in range(30):
            ops.table_generate(mode, lut, layout=layout, out=out)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(300):
                ops.table_generate(mode, lut, layout=layout, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 300)
        res.append("%s/%d %.4f ms %.2f TB/s" % (mode, layout, best, out.numel() * 2 / best / 1e9))
    ops.table_generate(mode, lut, layout=1, out=out)
    assert torch.equal(out.view(torch.int16), ref.view(torch.int16))
    del out, ref
print("  ".join(res))

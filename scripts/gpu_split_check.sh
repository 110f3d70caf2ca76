# parity + timing of the split generator on the GPU box
set -x
python -m pytest tests/test_gpu_tables.py -x -q 2>&1 | tail -5
python - <<'PY'
import torch, numpy as np, time
from iivision_b200 import ops
from iivision_b200._lib import ALGO_TREE, ALGO_SPLIT
from iivision_b200 import make_data_tables, palette
lut = np.random.default_rng(0).integers(0, 100, (16, 16)).astype(np.int32); lut = np.minimum(lut, lut.T); np.fill_diagonal(lut, 0)
for mode in ("HGR", "DHGR"):
    out = torch.empty(ops.table_shape(mode), dtype=torch.uint16, device="cuda")
    for layout in (1, 0):
        for name, algo in (("tree", ALGO_TREE), ("split", ALGO_SPLIT)):
            for _ in range(20): ops.table_generate(mode, lut, layout=layout, out=out, algo=algo)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200): ops.table_generate(mode, lut, layout=layout, out=out, algo=algo)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 200
            print(mode, "layout", layout, name, "%.4f ms" % ms, "%.2f TB/s" % (out.numel() * 2 / ms / 1e9))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r02g_split_launches.csv python - <<'PY' > /dev/null 2>&1
import torch, numpy as np
from iivision_b200 import ops
lut = np.random.default_rng(0).integers(0, 100, (16, 16)).astype(np.int32); lut = np.minimum(lut, lut.T); np.fill_diagonal(lut, 0)
out = torch.empty(ops.table_shape("HGR"), dtype=torch.uint16, device="cuda")
for _ in range(8): ops.table_generate("HGR", lut, out=out)
torch.cuda.synchronize()
PY
grep -v '^==' gpurun_out/r02g_split_launches.csv | tail -12

"""NVLink traffic of the fused generate + scatter kernel (iiv_table_generate_scatter), for ncu.

One process drives two GPUs: GPU 0 generates its half of the HGR table and stores every
finished run into BOTH tables (its own and GPU 1's, peer-mapped), exactly what rank 0 of a
two-rank parallel.generate_sharded_fused does -- without a second process whose kernels an
ncu replay would fall out of step with.

    gpurun --gpus 2 -- ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum \
        -k regex:tree_kernel --csv --log-file gpurun_out/r02_nvlink_scatter.csv \
        python scripts/diag_nvlink_scatter.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iivision_b200 import ops, palette  # noqa: E402


def main():
    assert torch.cuda.device_count() >= 2, "needs two GPUs"
    cudart = ctypes.CDLL("libcudart.so")
    for a, b in ((0, 1), (1, 0)):
        torch.cuda.set_device(a)
        assert torch.cuda.can_device_access_peer(a, b)
        rc = cudart.cudaDeviceEnablePeerAccess(b, 0)
        assert rc in (0, 704), rc          # 704: already enabled
    torch.cuda.set_device(0)
    lut = ops.lut_cie2000(palette.NTSCPalette.rgb_by_value())
    shape = ops.table_shape("HGR")
    tables = [torch.zeros(shape, dtype=torch.uint16, device="cuda:%d" % d) for d in (0, 1)]
    ptrs = [t.data_ptr() for t in tables]
    n = 1 << 14
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(3):
        ev[0].record()
        ops.table_generate_scatter("HGR", lut, ptrs, 0, 0, n // 2, layout=ops.LAYOUT_SYMMETRIC)
        ev[1].record()
        torch.cuda.synchronize()
        print("scatter of rows [0, %d): %.3f ms" % (n // 2, ev[0].elapsed_time(ev[1])))
    want = ops.table_generate("HGR", lut, layout=ops.LAYOUT_SYMMETRIC, row_end=n // 2)
    torch.cuda.synchronize()
    half = shape[1] // 2
    ok0 = torch.equal(tables[0][:, :half], want[:, :half])
    ok1 = torch.equal(tables[1][:, :half].to("cuda:0"), want[:, :half])
    print("rows landed in both tables: local %s, peer %s" % (ok0, ok1))
    print("bytes stored to the peer per launch: %d" % (2 * half * 2))


if __name__ == "__main__":
    main()

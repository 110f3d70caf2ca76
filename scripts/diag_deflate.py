"""Device-side deflate of a table (csrc/iiv_deflate.cu): stage times, size against the host
zlib -6 writer (npz_io.savez_compressed), and make_edit_distance wall clock.

    gpurun -- python scripts/diag_deflate.py [HGR|DHGR] > gpurun_out/deflate.txt
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iivision_b200 import colours, make_data_tables as mdt, npz_io, ops, palette, screen  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "HGR"
    bitmap_cls = screen.HGRBitmap if mode == "HGR" else screen.DHGRBitmap
    nominal = colours.HGRColours if mode == "HGR" else colours.DHGRColours
    pal = palette.NTSCPalette
    edp = mdt.compute_substitute_costs(pal)
    table = mdt.compute_edit_distance_device(edp, bitmap_cls, ops.LAYOUT_TRIANGULAR)
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        stream, sizes, crcs, block_bytes = ops.deflate_table(mode, table)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("%s deflate_table call %d: %.1f ms, %d -> %d bytes (ratio %.3f)"
              % (mode, rep, 1e3 * dt, table.numel() * 2, stream.numel(),
                 table.numel() * 2 / stream.numel()))
    timings = {}
    ops.deflate_table(mode, table, timings=timings)
    for name, ms in timings.items():
        print("    %-28s %8.3f ms" % (name, ms))
    with tempfile.TemporaryDirectory() as tmp:
        old = mdt.DATA_DIR
        mdt.DATA_DIR = tmp
        try:
            for rep in range(3):
                t0 = time.perf_counter()
                mdt.make_edit_distance(pal, edp, bitmap_cls, nominal)
                dt = time.perf_counter() - t0
                path = os.path.join(tmp, "%s_palette_%d_edit_distance.npz"
                                    % (bitmap_cls.NAME, pal.ID.value))
                print("make_edit_distance %d: %.3f s, file %d bytes"
                      % (rep, dt, os.path.getsize(path)))
        finally:
            mdt.DATA_DIR = old
        ours = os.path.getsize(path)
        t0 = time.perf_counter()
        with np.load(path) as z:
            back = z["edit_distance"]
        print("np.load of it: %.2f s" % (time.perf_counter() - t0))
        host = table.cpu().numpy().view(np.uint16)
        assert np.array_equal(back, host)
        t0 = time.perf_counter()
        got = npz_io.load_member(path, "edit_distance")
        print("npz_io.load_member (parallel inflate): %.2f s" % (time.perf_counter() - t0))
        assert np.array_equal(got, host)
        ref = os.path.join(tmp, "host.npz")
        t0 = time.perf_counter()
        npz_io.savez_compressed(ref, edit_distance=host)
        print("host zlib -6 writer (npz_io.savez_compressed, %d threads): %.2f s, file %d bytes"
              % (os.cpu_count(), time.perf_counter() - t0, os.path.getsize(ref)))
        print("size ratio ours / zlib-6 = %.3f" % (ours / os.path.getsize(ref)))


if __name__ == "__main__":
    main()

"""Where the time of the tables' e2e step (compute_substitute_costs + compute_edit_distance ->
host array) goes: cProfile over five calls after a warm-up."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from iivision_b200 import colours, make_data_tables, palette, screen  # noqa: E402


def step():
    edp = make_data_tables.compute_substitute_costs(palette.NTSCPalette)
    return make_data_tables.compute_edit_distance(edp, screen.HGRBitmap, colours.HGRColours)


step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
res = None
for _ in range(5):
    res = None
    t0 = time.perf_counter()
    res = step()
    print("step %.1f ms" % ((time.perf_counter() - t0) * 1e3))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

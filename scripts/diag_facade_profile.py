"""cProfile of the video.Video facade under the Movie.encode schedule (bench.facade_figures)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

print(bench.facade_figures(torch, n_frames=4))       # warm: tables, allocator, kernels
pr = cProfile.Profile()
pr.enable()
res = bench.facade_figures(torch, n_frames=8)
pr.disable()
print(res)
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)

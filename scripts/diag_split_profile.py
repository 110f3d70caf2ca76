"""One HGR table through the split generator, for ncu captures of split_prologue /
split_kernel (scripts/gpu_split_profile.sh)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iivision_b200 import ops  # noqa: E402

lut = np.random.default_rng(0).integers(0, 100, (16, 16)).astype(np.int32)
lut = np.minimum(lut, lut.T)
np.fill_diagonal(lut, 0)
out = torch.empty(ops.table_shape("HGR"), dtype=torch.uint16, device="cuda")
for _ in range(6):
    ops.table_generate("HGR", lut, out=out)
torch.cuda.synchronize()

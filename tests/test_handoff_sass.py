"""CPU: the built encoder's cross-warp hand-offs, at the SASS level (scripts/check_handoffs.py).
Guards against the defect found in round 1 -- ptxas moving a data load out of the spin loop
on its flag -- without needing a GPU."""

import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not (shutil.which("cuobjdump") and shutil.which("nvdisasm")),
                    reason="CUDA binary utilities not on PATH")
def test_record_poll_reloads_and_block_wait_fences():
    from iivision_b200 import _build
    _build.build()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_handoffs.py")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert r.stdout.count("LDS.128") >= 4

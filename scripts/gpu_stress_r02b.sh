# repeated fresh-process runs of the late round-2 kernels' parity tests (split generator with
# its dependent launches, factored scorer)
fails=0
for k in $(seq 1 ${RUNS:-10}); do
  python -m pytest tests/test_gpu_tables.py tests/test_gpu_factored.py -x -q 2>&1 | tail -1 | grep -q "passed" || fails=$((fails+1))
done
echo "stress: $fails failures in ${RUNS:-10} fresh processes"
python - <<'PY'
import numpy as np, torch
from iivision_b200 import ops
# many back-to-back generates with alternating LUTs / modes / layouts / row ranges, checked at the end
rng = np.random.default_rng(0)
luts = []
for _ in range(4):
    l = rng.integers(0, 200, (16, 16)).astype(np.int32); l = np.minimum(l, l.T); np.fill_diagonal(l, 0); luts.append(l)
want = {(m, k): ops.table_generate(m, luts[k], algo=1) for m in ("DHGR",) for k in range(4)}
out = {k: torch.empty_like(want[("DHGR", 0)]) for k in range(4)}
bad = 0
for rep in range(200):
    k = rep % 4
    ops.table_generate("DHGR", luts[k], out=out[k])
    if rep % 7 == 0:
        ops.table_generate("DHGR", luts[(k + 1) % 4], out=out[(k + 1) % 4], row_begin=100, row_end=5000)
        ops.table_generate("DHGR", luts[(k + 1) % 4], out=out[(k + 1) % 4])
torch.cuda.synchronize()
for k in range(4):
    bad += int(not torch.equal(out[k].view(torch.int16), want[("DHGR", k)].view(torch.int16)))
print("back-to-back generates with alternating LUTs: %d of 4 tables wrong" % bad)
PY

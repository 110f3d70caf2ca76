set -x
python -m pytest tests/test_gpu_factored.py tests/test_long_streams.py -x -q -m gpu 2>&1 | tail -5
python bench.py --scorer-only 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d.get('scorer',d)
for k in ('scored_frames_per_s','scored_frames_per_s_5pct_change','scored_frames_per_s_table_gathers','scored_frames_per_s_table_gathers_5pct_change','factored_equals_table_path'): print(k, s.get(k))
r=s['roofline']; print({k:v for k,v in r.items() if k not in ('note','table_gather_kernel')})
"

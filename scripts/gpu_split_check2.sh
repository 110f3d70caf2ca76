python -m pytest tests/test_gpu_tables.py tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python scripts/diag_split_time.py

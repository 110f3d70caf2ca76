set -x
IIV_RANDOM_CASES=40 timeout 400 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_encoder_random.py tests/test_gpu_golden.py tests/test_gpu_facade.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --scorer-only 2>&1 | tail -1

timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
for n in 1 1 1 64 148 148 148 296; do
  echo -n "clips=$n: "; IIV_BENCH_CLIPS=$n python bench.py --scorer-only 2>&1 | tail -1 | grep -o "single-clip[^\"]*"
done

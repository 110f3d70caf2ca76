set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; tail -c 2600 gpurun_out/bench_v3.json; tail -5 gpurun_out/bench_v3.err

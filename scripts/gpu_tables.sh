set -x
python -m pytest tests/test_gpu_tables.py tests/test_gpu_golden.py -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 --no-scorer --no-cpu-baseline > gpurun_out/bench_tree.json 2> gpurun_out/bench_tree.err; tail -c 1500 gpurun_out/bench_tree.json; tail -5 gpurun_out/bench_tree.err
ncu --set full --clock-control none --import-source on -k regex:tree_kernel -s 3 -c 1 -o gpurun_out/prof_tree python bench.py --steps 2 --warmup 1 --no-scorer --no-cpu-baseline > gpurun_out/ncu_tree.log 2>&1

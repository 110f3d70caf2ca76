set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python __graft_entry__.py --smoke 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; tail -c 3000 gpurun_out/bench_v1.json; tail -5 gpurun_out/bench_v1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v1.csv python bench.py --steps 2 --warmup 1 --no-scorer --no-cpu-baseline > gpurun_out/ncu_launch_v1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 3 -c 1 -o gpurun_out/prof_chain_v1 python bench.py --steps 2 --warmup 1 --no-scorer --no-cpu-baseline > gpurun_out/ncu_chain_v1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 1 -c 1 -o gpurun_out/prof_encode_v1 python bench.py --scorer-only > gpurun_out/ncu_encode_v1.log 2>&1
ls -la gpurun_out

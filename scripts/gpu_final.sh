set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_final.json
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 4000 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tree_kernel -s 3 -c 1 -o gpurun_out/prof_tree_final python bench.py --steps 2 --warmup 1 --no-scorer --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 2 -c 1 -o gpurun_out/prof_encode_final python bench.py --scorer-only > /dev/null 2>&1
ls -la gpurun_out | tail -8

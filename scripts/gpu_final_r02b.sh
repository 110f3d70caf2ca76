# round 2 (late) evidence after the split generator: tests, smoke, both bench arms, launch list,
# ncu captures of the generator's kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02h_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r02h_bench_reference.json
python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -c 1500 gpurun_out/r02h_bench.json; tail -3 gpurun_out/r02h_bench.err
IIV_BENCH_LONG_FRAMES=60 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02h_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:split_kernel -s 3 -c 1 -o gpurun_out/r02h_prof_split python bench.py --steps 2 --warmup 1 --no-scorer --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:split_prologue -s 3 -c 1 -o gpurun_out/r02h_prof_split_prologue python bench.py --steps 2 --warmup 1 --no-scorer --no-cpu-baseline > /dev/null 2>&1
IIV_BENCH_LONG_FRAMES=60 ncu --set full --clock-control none --import-source on -k regex:score_frames_factored -s 3 -c 1 -o gpurun_out/r02h_prof_factored python bench.py --scorer-only > /dev/null 2>&1
ls -la gpurun_out | tail -8

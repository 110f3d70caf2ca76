set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
./scripts/build/diag_hbm_fill > gpurun_out/r02_hbm_fill.txt 2>&1; cat gpurun_out/r02_hbm_fill.txt
python bench.py --scorer-only > gpurun_out/r02a_scorer.json 2> gpurun_out/r02a_scorer.err; tail -c 6000 gpurun_out/r02a_scorer.json; tail -5 gpurun_out/r02a_scorer.err
IIV_BENCH_LONG_FRAMES=60 ncu --set full --clock-control none --import-source on -k regex:score_frames -s 3 -c 1 -o gpurun_out/r02a_prof_score python bench.py --scorer-only > /dev/null 2>&1
IIV_BENCH_LONG_FRAMES=60 ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 9 -c 1 -o gpurun_out/r02a_prof_encode python bench.py --scorer-only > /dev/null 2>&1
ls -la gpurun_out | tail -6

IIV_BENCH_CLIPS=${PROFCLIPS:-148} ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 2 -c 1 -o gpurun_out/prof_encode_v7 python bench.py --scorer-only > gpurun_out/ncu_encode_v7.log 2>&1
tail -2 gpurun_out/ncu_encode_v7.log

set -x
ncu --set full --clock-control none --import-source on -k regex:score_frames_factored -s 3 -c 1 -o gpurun_out/r02i_prof_factored python bench.py --scorer-only > /dev/null 2>&1
ls -la gpurun_out/r02i*

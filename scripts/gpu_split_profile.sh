set -x
ncu --set full --clock-control none --import-source on -k regex:split_kernel -s 3 -c 1 -o gpurun_out/r02g_prof_split python scripts/diag_split_profile.py > gpurun_out/r02g_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:split_prologue -s 3 -c 1 -o gpurun_out/r02g_prof_split_prologue python scripts/diag_split_profile.py > gpurun_out/r02g_prof.log 2>&1
ls -la gpurun_out/r02g_prof_split*

# round 2: memcheck + racecheck + stress of the encoder after the move to acquire/release hand-offs
mkdir -p gpurun_out
bash scripts/gpu_sanitize.sh > gpurun_out/r02_memcheck.txt 2>&1; tail -4 gpurun_out/r02_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_encoder.py::test_out_of_work_and_requeue -x -q > gpurun_out/r02_racecheck.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r02_racecheck.txt | sort | uniq -c | sort -rn | head -12
grep -E "iiv_encoder.cu:[0-9]+" -o gpurun_out/r02_racecheck.txt | sort | uniq -c | sort -rn | head -40 > gpurun_out/r02_racecheck_lines.txt; cat gpurun_out/r02_racecheck_lines.txt
RUNS=${RUNS:-16} bash scripts/gpu_stress.sh | tail -3

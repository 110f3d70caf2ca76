which compute-sanitizer || ls /usr/local/cuda/bin | grep -i sanit
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_encoder.py::test_out_of_work_and_requeue tests/test_gpu_scorer.py tests/test_gpu_tables.py::test_row_ranges_unaligned tests/test_gpu_tables.py::test_string_distance -x -q 2>&1 | tail -8
echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_encoder.py::test_out_of_work_and_requeue" -x -q 2>&1 | tail -12
echo "racecheck rc=$?"

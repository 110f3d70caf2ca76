# memcheck of the kernels added late in round 2: split generator (row ranges, layouts, extreme
# LUTs, two streams) and the factored scorer
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest \
  tests/test_gpu_tables.py::test_row_ranges_unaligned tests/test_gpu_tables.py::test_row_blocks_compose \
  "tests/test_gpu_tables.py::test_tree_kernel_equals_chain_kernel[0-DHGR]" \
  tests/test_gpu_tables.py::test_concurrent_generates_on_two_streams \
  "tests/test_gpu_factored.py::test_factored_extreme_luts[HGR]" \
  "tests/test_gpu_factored.py::test_factored_equals_table_path[5-DHGR]" \
  "tests/test_gpu_factored.py::test_factored_equals_table_path[0-HGR]" \
  tests/test_gpu_factored.py::test_factored_bad_arguments -x -q > gpurun_out/r02b_memcheck.txt 2>&1
echo "exit $?" >> gpurun_out/r02b_memcheck.txt
tail -6 gpurun_out/r02b_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest \
  "tests/test_gpu_factored.py::test_factored_equals_table_path[5-DHGR]" -x -q > gpurun_out/r02b_racecheck.txt 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r02b_racecheck.txt | sort | uniq -c | sort -rn | head -8

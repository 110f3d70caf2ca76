fail=0
for i in $(seq 1 ${RUNS:-24}); do
  timeout 200 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_golden.py::test_opcode_streams -x -q > /tmp/stress_$i.log 2>&1
  tail -1 /tmp/stress_$i.log | grep -q passed || { fail=$((fail+1)); echo "run $i FAILED"; grep -E "^E |Error|first differing|assert" /tmp/stress_$i.log | head -12; grep -E "^(FAILED|tests/)" /tmp/stress_$i.log | head -3; }
done
echo "failures: $fail"

# builds encoder variants on the GPU box (IIV_NVCC_FLAGS) and times each: usage
#   gpurun -- bash scripts/gpu_variants.sh "" "-DIIV_X_A" "-DIIV_X_A -DIIV_X_B"
for flags in "$@"; do
  echo "=== variant: [$flags]"
  touch iivision_b200/csrc/iiv_encoder.cu
  IIV_NVCC_FLAGS="$flags" python -m iivision_b200._build > /dev/null || exit 1
  timeout 300 python bench.py --scorer-only 2>/dev/null | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())['scorer']
t = d['encoded_trace_clip0']; s = d['single_clip_trace']
print('148 clips %.0f fps; single %.1f fps (A %d loop %d wait_rows %d wait_mt %d) long %.1f hgr60 %.1f ok %s' % (
    d['encoded_frames_per_s'], d['single_clip_frames_per_s'], s['cycles_score_heapify'], s['cycles_opcode_loop'],
    s['cycles_wait_rows'], s['cycles_wait_mt_applier'], d['config3_long_clip']['frames_per_s'], d['hgr_60_frame_clip_frames_per_s'], d['config3_long_clip']['matches_reference']))"
done
touch iivision_b200/csrc/iiv_encoder.cu
if [ -f scripts/tmp_old_encoder.cu.txt ]; then
  echo "=== previous commit's encoder"
  cp iivision_b200/csrc/iiv_encoder.cu /tmp/new_encoder.cu
  cp scripts/tmp_old_encoder.cu.txt iivision_b200/csrc/iiv_encoder.cu
  python -m iivision_b200._build > /dev/null || exit 1
  timeout 300 python bench.py --scorer-only 2>/dev/null | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())['scorer']
s = d['single_clip_trace']
print('148 clips %.0f fps; single %.1f fps (A %d loop %d wait_rows %d wait_mt %d) long %.1f' % (
    d['encoded_frames_per_s'], d['single_clip_frames_per_s'], s['cycles_score_heapify'], s['cycles_opcode_loop'],
    s['cycles_wait_rows'], s['cycles_wait_mt_applier'], d['config3_long_clip']['frames_per_s']))"
  cp /tmp/new_encoder.cu iivision_b200/csrc/iiv_encoder.cu
fi

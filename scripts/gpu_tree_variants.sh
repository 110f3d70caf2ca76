# builds tree_kernel variants on the GPU box and times the headline step: usage
#   gpurun -- bash scripts/gpu_tree_variants.sh "" "-DIIV_TREE_MIN_BLOCKS=6"
for flags in "$@"; do
  echo "=== variant: [$flags]"
  touch iivision_b200/csrc/iiv_tables.cu
  IIV_NVCC_FLAGS="$flags" python -m iivision_b200._build > /dev/null || exit 1
  python bench.py --no-scorer --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())
print('ms_per_step %.4f  kernel_ms %.4f  frac %.3f  min/max %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['step_ms_min_max']))"
done
touch iivision_b200/csrc/iiv_tables.cu

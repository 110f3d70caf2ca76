# builds split_kernel variants on the GPU box and times the generator's step: usage
#   gpurun -- bash scripts/gpu_split_variants.sh "" "-DIIV_SPLIT_EPT=8 -DIIV_SPLIT_MIN_BLOCKS=8"
for flags in "$@"; do
  echo "=== variant: [$flags]"
  touch iivision_b200/csrc/iiv_tables.cu
  IIV_NVCC_FLAGS="$flags" python -m iivision_b200._build > /dev/null || exit 1
  python scripts/diag_split_time.py
done
touch iivision_b200/csrc/iiv_tables.cu

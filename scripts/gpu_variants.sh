for mb in 4 5 6 8; do
  echo -n "mb=$mb: "; IIV_LIB_PATH=$PWD/iivision_b200/libiiv_mb$mb.so python bench.py --steps 50 --warmup 3 --no-scorer --no-cpu-baseline 2>&1 | tail -1 | grep -oE '"ms_per_step": [0-9.]+|"kernel_ms": [0-9.]+' | tr '\n' ' '; echo
done

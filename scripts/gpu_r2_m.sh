for v in lib lib_bk5; do
  echo "== $v"; MCU_PLAN_TIMING=1 MORPHO_CUDA_LIB=/root/repo/morpho_b200/$v/libmorpho_cuda.so python scripts/block_case.py 2>&1 | tail -2
done

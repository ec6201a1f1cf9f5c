for v in lib lib_w14c2 lib_w12c2 lib_w20c1 lib_w24c1; do
  extra=""
  case $v in *c1) extra="--opt patch_smem_per_cta=200000 --opt patch_stages=4";; esac
  MORPHO_CUDA_LIB=/root/repo/morpho_b200/$v/libmorpho_cuda.so timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --quick $extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['config']['kernel_ms'], d['roofline']['frac'])"
done

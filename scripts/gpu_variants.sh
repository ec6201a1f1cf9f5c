# times bench.py --quick with alternative builds of the patch kernel (morpho_b200/lib_<variant>/, see csrc/Makefile VARIANT=)
for v in lib lib_w12p4 lib_w16p4; do
  extra=""
  case $v in lib_w12p4) extra="--patch-owned 384";; esac
  [ -f morpho_b200/$v/libmorpho_cuda.so ] || continue
  MORPHO_CUDA_LIB=/root/repo/morpho_b200/$v/libmorpho_cuda.so timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --quick $extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['config']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done

python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for cfg in "patch_tiles=0" "patch_tiles=0 --opt patch_group=16" "patch_tiles=0 --opt patch_max_owned=480"; do
  timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --quick --opt $cfg 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', d['value'], d['config']['kernel_ms'], d['roofline']['frac'])"
done

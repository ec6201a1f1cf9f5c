python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for cfg in "patch_tiles=0" "patch_tiles=1" "patch_tiles=1 --opt patch_stages=2" "patch_tiles=0 --opt patch_group=1"; do
  timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --quick --opt $cfg 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', d['value'], d['config']['kernel_ms'], d['roofline']['frac'])"
done
timeout 300 python scripts/patch_timeline.py 707 | grep -E "alive|by CTA"
timeout 300 python scripts/patch_trace.py | tail -12

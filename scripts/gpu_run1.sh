set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench7.json')); print(d['value'], d['config']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"
tail -3 gpurun_out/bench7.err

set -x
( time python bench.py --steps 200 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err ); echo "bench rc=$?"; tail -5 gpurun_out/bench_r2c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2c.json')); print(d['value'], d['e2e']); print(d['roofline']['frac'])"

# multi-GPU check (gpurun --gpus N): [the world-size-2 tests,] then the bench at N as the driver launches it
N=${1:-2}
set -x
if [ "$2" = "tests" ]; then
( time timeout 400 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/pytest_multigpu.log 2>&1 ); echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multigpu.log
fi
timeout ${3:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench$N rc=$?"
grep "phase\|strong\|configs\|parity\|Error\|error\|watch" gpurun_out/bench_n$N.err | cut -c1-2500 | tail -20
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read())
print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['kernel_ms'], d['e2e']['value'], d['clocks'])
PY

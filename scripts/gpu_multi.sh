# multi-GPU check (gpurun --gpus N): the world-size-2 tests, then the bench at N as the driver launches it
N=${1:-2}
set -x
( time python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/pytest_multigpu.log 2>&1 ); echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multigpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench$N rc=$?"
grep -v "^\[rank [1-9]" gpurun_out/bench_n$N.err | tail -8
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read())
print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['kernel_ms'], d['e2e']['value'], d['clocks'])
print(d.get('strong_scaling'))
PY
if [ "$2" = "nccl" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 --exchange nccl --quick --scaling weak > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "bench$N nccl rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 5 --exchange serial --quick --scaling weak > gpurun_out/bench_n${N}_serial.json 2> gpurun_out/bench_n${N}_serial.err; echo "bench$N serial rc=$?"
python - <<PY
import json
for k in ("nccl","serial"):
    d=json.loads(open("gpurun_out/bench_n${N}_%s.json" % k).read())
    print(k, d['n_gpus'], d['value'], d['ms_per_step'], d['config']['kernel_ms'])
PY
fi

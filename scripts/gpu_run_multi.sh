set -x
nvidia-smi -L
python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['kernel_ms'], d['e2e']['value'])"
python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline --quick 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['kernel_ms'])"

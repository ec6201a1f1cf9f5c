N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench$N rc=$?"
tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['kernel_ms'], d['e2e']['value'], d['clocks'])"

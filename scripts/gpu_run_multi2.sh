python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -5
bash scripts/gpu_run_multi.sh 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 5 --no-fused-push --quick 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no-fused', d['n_gpus'], d['value'], d['ms_per_step'], d['config']['kernel_ms'])"

set -x
nvidia-smi -L
( time python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2a.log 2>&1 ); echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r2a.log
( time python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err ); echo "bench rc=$?"; tail -12 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json | head -c 3000

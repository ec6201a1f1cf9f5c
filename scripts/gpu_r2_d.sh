set -x
nvidia-smi -L
( time python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu_r2d.log 2>&1 ); echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r2d.log
( time python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err ); echo "bench rc=$?"; tail -12 gpurun_out/bench_r2d.err; head -c 1500 gpurun_out/bench_r2d.json

set -x
( time python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2e.log 2>&1 ); echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_r2e.log

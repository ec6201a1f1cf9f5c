set -x
( time python -m pytest tests/test_extension.py -m gpu -q -s -k "pairwise or thomson" > gpurun_out/pytest_gpu_r2f.log 2>&1 ); echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu_r2f.log

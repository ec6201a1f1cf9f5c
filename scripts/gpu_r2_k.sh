set -x
( time python -m pytest tests/test_parity_gpu.py -m gpu -q -k "block" > gpurun_out/pytest_gpu_r2k.log 2>&1 ); echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2k.log
python scripts/block_case.py 2>&1 | tee gpurun_out/block_case.log

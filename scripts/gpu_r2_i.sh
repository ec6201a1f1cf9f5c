set -x
( time python -m pytest tests/test_parity_gpu.py -m gpu -q -s -k "block or golden" > gpurun_out/pytest_gpu_r2i.log 2>&1 ); echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r2i.log
MCU_PLAN_TIMING=1 python scripts/block_bench.py > gpurun_out/block_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/block_bench.log

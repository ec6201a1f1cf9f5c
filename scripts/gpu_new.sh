set -x
( time python -m pytest tests/test_parity_gpu.py tests/test_extension.py -m gpu -x -q -k "lowering or connectivity or volume_integral" > gpurun_out/pytest_new.log 2>&1 ); echo "pytest rc=$?"; tail -30 gpurun_out/pytest_new.log | cut -c1-1500

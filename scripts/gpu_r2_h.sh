set -x
( time python -m pytest tests/test_extension.py -m gpu -q -s -k "tactoid or field_functionals" > gpurun_out/pytest_gpu_r2h.log 2>&1 ); echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu_r2h.log

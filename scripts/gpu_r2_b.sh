set -x
( time python -m pytest tests/test_extension.py -m gpu -q -s -x > gpurun_out/pytest_ext_r2b.log 2>&1 ); echo "pytest rc=$?"; tail -40 gpurun_out/pytest_ext_r2b.log

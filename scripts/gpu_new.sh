set -x
( time python -m pytest tests/test_extension.py -m gpu -x -q -k "connectivity" > gpurun_out/pytest_new.log 2>&1 ); echo "pytest rc=$?"; tail -25 gpurun_out/pytest_new.log | cut -c1-1200

set -x
( time python -m pytest tests/test_integral.py tests/test_extension.py -m gpu -x -q -k "not reference_suite and not tactoid and not catenoid_example" > gpurun_out/pytest_new.log 2>&1 ); echo "pytest rc=$?"; tail -30 gpurun_out/pytest_new.log | cut -c1-1800

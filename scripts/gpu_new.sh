set -x
( time python -m pytest tests/test_extension.py -m gpu -x -q -s -k "more_examples and (electrostatics or cube)" > gpurun_out/pytest_new.log 2>&1 ); echo "pytest rc=$?"; grep "as shipped\|passed\|failed\|Error\|assert" gpurun_out/pytest_new.log | cut -c1-600 | tail -20

python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_run_multi.sh 2

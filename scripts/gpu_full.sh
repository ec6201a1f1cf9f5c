# full GPU check of the tree: every -m gpu test, smoke(), then the bench as the driver runs it
set -x
( time python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_full.log 2>&1 ); echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-600

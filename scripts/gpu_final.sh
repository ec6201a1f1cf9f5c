set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 500 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"
tail -3 gpurun_out/bench_final_ref.err

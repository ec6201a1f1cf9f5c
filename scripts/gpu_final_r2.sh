# final check of the tree on one B200: every -m gpu test, smoke(), the bench as the driver runs it, the reference arm,
# then (only after the plain runs exited 0) the ncu captures of the headline kernel
set -x
( time python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1 ); echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_final.err | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
sha256sum morpho_b200/csrc/mcu_patch.cu | cut -c1-16 > gpurun_out/patch_sha16.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --quick --no-parity --no-configs > gpurun_out/plain.log 2>&1 &&
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_patch_ring -s 6 -c 2 -f -o gpurun_out/prof_ring_r2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --quick --no-parity --no-configs > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_ring_r2.ncu-rep > gpurun_out/prof_ring_r2_summary.txt 2>&1; head -40 gpurun_out/prof_ring_r2_summary.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --quick --no-parity --no-configs > gpurun_out/plain2.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 60 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --quick --no-parity --no-configs > gpurun_out/ncu_list.log 2>&1
echo "ncu launches rc=$?"

set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --quick > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_patch_ring -s 6 -c 2 -f -o gpurun_out/prof_ring_r1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu7.log 2>&1
echo "ncu rc=$?"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --quick > gpurun_out/plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu8.log 2>&1
echo "ncu launches rc=$?"

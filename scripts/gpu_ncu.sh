set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_patch_ring -s 6 -c 2 -f -o gpurun_out/prof_ring_r1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu7.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu7.log

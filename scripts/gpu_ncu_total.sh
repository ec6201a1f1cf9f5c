# ncu --set full of the patch TOTAL kernel (triangle lists) on the headline mesh
python scripts/maps_bench.py 707 > gpurun_out/plain_maps2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_patch_ring -s 2 -c 1 -f -o gpurun_out/prof_total_r1 python scripts/maps_bench.py 707 > gpurun_out/ncu_maps2.log 2>&1
echo "ncu rc=$?"

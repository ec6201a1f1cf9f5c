set -x
python scripts/pairwise_bench.py 200000 2 > gpurun_out/plain_pw.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pairwise_coulomb -s 2 -c 2 -f -o gpurun_out/prof_pairwise_r1 python scripts/pairwise_bench.py 200000 2 > gpurun_out/ncu_pw.log 2>&1
echo "ncu pairwise rc=$?"
python scripts/maps_bench.py 707 > gpurun_out/plain_maps.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_patch_ring|k_integrand" -s 12 -c 6 -f -o gpurun_out/prof_maps_r1 python scripts/maps_bench.py 707 > gpurun_out/ncu_maps.log 2>&1
echo "ncu maps rc=$?"

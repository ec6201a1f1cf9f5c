# ncu --set full of the all-pairs gradient kernel at 200 000 points (BASELINE config 3)
python scripts/pairwise_multi.py > gpurun_out/plain_pw2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pairwise_coulomb -s 2 -c 1 -f -o gpurun_out/prof_pairwise_r1b python scripts/pairwise_multi.py > gpurun_out/ncu_pw2.log 2>&1
echo "ncu rc=$?"

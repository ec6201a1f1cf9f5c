set -x
python scripts/block_case.py > gpurun_out/block_case.log 2>&1 && cat gpurun_out/block_case.log &&
ncu --set full --clock-control none --import-source on -k regex:k_block_grad -s 3 -c 1 -f -o gpurun_out/prof_block_r2 python scripts/block_case.py > gpurun_out/ncu_block.log 2>&1
echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_block_r2.ncu-rep > gpurun_out/prof_block_r2_summary.txt 2>&1; cat gpurun_out/prof_block_r2_summary.txt

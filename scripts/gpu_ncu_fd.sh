# ncu --set full of the finite-difference kernels of config 4 (Nematic gradient / fieldgradient on a 1M-triangle disk)
python scripts/field_bench.py 408 > gpurun_out/plain_fd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_fd_grad|k_fd_elem" -c 2 -f -o gpurun_out/prof_fd_r1 python scripts/field_bench.py 408 > gpurun_out/ncu_fd.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/plain_fd.log

CS=4096 python scripts/gpu_kernel_time.py > gpurun_out/plain_kt2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_logistic_tc -s 6 -c 2 -o gpurun_out/prof_r2a -f env CS=4096 python scripts/gpu_kernel_time.py > gpurun_out/ncu_kt2.log 2>&1
tail -3 gpurun_out/plain_kt2.log gpurun_out/ncu_kt2.log

CMD="python scripts/gpu_kernel_time.py"
export CS=4096 REF=1
timeout 300 $CMD > gpurun_out/plain_kt3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_logistic_tc -s 8 -c 1 -o gpurun_out/prof_r2b -f $CMD > gpurun_out/ncu_kt3.log 2>&1
tail -2 gpurun_out/plain_kt3.log gpurun_out/ncu_kt3.log
unset CS REF
WHICH=funnel,iid timeout 600 python scripts/gpu_secondary.py 2>&1 | tee -a gpurun_out/secondary_r2.jsonl | cut -c1-400

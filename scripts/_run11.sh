CMD="python scripts/gpu_advance_probe.py"
timeout 300 $CMD > gpurun_out/plain_adv.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advance -s 1 -c 1 -o gpurun_out/prof_adv -f $CMD > gpurun_out/ncu_adv.log 2>&1
cat gpurun_out/plain_adv.log | tail -n 2

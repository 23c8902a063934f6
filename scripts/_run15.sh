CMD="python bench.py --steps 1 --warmup 3 --adapt 3 --transitions 2 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_final.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_l_final.log 2>&1
tail -n 2 gpurun_out/ncu_l_final.log | cut -c1-300

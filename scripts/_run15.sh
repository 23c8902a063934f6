CMD="python bench.py --chains 512 --steps 1 --warmup 3 --adapt 3 --transitions 2 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_512.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/launches_512.csv $CMD > gpurun_out/ncu_l_512.log 2>&1
tail -c 600 gpurun_out/plain_512.log

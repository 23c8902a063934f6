set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_pytest.log
timeout 900 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 3000 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
CMD="python bench.py --steps 1 --warmup 3 --adapt 3 --transitions 2 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_r2b.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_l2.log 2>&1
tail -2 gpurun_out/ncu_l2.log

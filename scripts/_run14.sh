set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.log
timeout 1200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 2800 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
CMD="python bench.py --steps 1 --warmup 3 --adapt 3 --transitions 2 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_final.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_l_final.log 2>&1
export CS=4096 REF=1
CMD2="python scripts/gpu_kernel_time.py"
timeout 300 $CMD2 > gpurun_out/plain_kt_final.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_logistic_tc -s 8 -c 1 -o gpurun_out/prof_final -f $CMD2 > gpurun_out/ncu_kt_final.log 2>&1
tail -n 1 gpurun_out/plain_kt_final.log

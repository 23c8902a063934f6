timeout 1200 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 2500 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err

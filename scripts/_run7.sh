timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_pytest.log
timeout 1200 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 2600 gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err

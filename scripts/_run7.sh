timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps 2 --no-cpu-baseline > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; python -c "
import json
r=json.loads(open('gpurun_out/bench_r2e.json').read().strip().splitlines()[-1])
print(r['n_gpus'], round(r['value']), round(r['e2e']['value']), r['roofline']['kernel_share_of_step'], r['roofline']['avg_launch_ms'], r['clocks'])
"; tail -3 gpurun_out/bench_r2e.err

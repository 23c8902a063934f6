timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gauss" 2>&1 | tail -25 | tee gpurun_out/r2_pytest_gauss.log
WHICH=gauss timeout 600 python scripts/gpu_secondary.py 2>&1 | tee gpurun_out/secondary_r2c.jsonl | cut -c1-420

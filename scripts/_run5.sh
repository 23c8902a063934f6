timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_dense_metric.py -m gpu -x -q -k "gauss or dense or odd" 2>&1 | tail -3
WHICH=gauss timeout 600 python scripts/gpu_secondary.py 2>&1 | tee gpurun_out/secondary_r2.jsonl | cut -c1-700

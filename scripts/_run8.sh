timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_pytest.log
timeout 600 python scripts/gpu_secondary.py 2>&1 | tee gpurun_out/secondary_r2b.jsonl | cut -c1-330

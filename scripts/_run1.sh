for ref in 0 1; do
  REF=$ref CS=4096,512 python scripts/gpu_kernel_time.py 2>&1 | tail -2 | tee -a gpurun_out/r2_ktime.log
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -5

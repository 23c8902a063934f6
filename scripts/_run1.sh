for ref in 0 1; do
BNUTS_LIB=build/libbnuts_g2o1.so REF=$ref CS=4096 timeout 90 python scripts/gpu_kernel_time.py 2>&1 | tail -1 | tee -a gpurun_out/r2_ktime.log
done

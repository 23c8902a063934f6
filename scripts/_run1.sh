for rep in 1 2; do
for lib in "" build/libbnuts_pf0.so; do
BNUTS_LIB=$lib REF=1 CS=4096,512 timeout 120 python scripts/gpu_kernel_time.py 2>&1 | tail -2
done; done

for v in g1r0x00 g1r0x55 g2r0x00 g2r0x55; do
BNUTS_LIB=build/libbnuts_$v.so REF=1 CS=4096,128 timeout 120 python scripts/gpu_kernel_time.py 2>&1 | tail -2
done

# timing of the ablated builds of logistic_rm.cu (scripts/build_rm_variant.sh rmd<k> -DBNUTS_RM_DEBUG=<k>)
export REF=1 CS=4096 REF_FILE=gpurun_out/ref_b.npy
timeout -s KILL 120 python scripts/gpu_kernel_time.py 2>&1 | grep -v Warn
export BNUTS_TC_DEBUG_NOCHECK=1
for t in 1 4 8 6 14 15; do echo "RM_DEBUG=$t"; BNUTS_LIB=build/libbnuts_rmd$t.so timeout -s KILL 120 python scripts/gpu_kernel_time.py 2>&1 | grep -v Warn | tail -3; done

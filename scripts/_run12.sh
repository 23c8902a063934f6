for rep in 1 2; do
for lib in "" build/libbnuts_inl.so; do
echo "== lib=$lib"
BNUTS_LIB=$lib WHICH=funnel,iid timeout 600 python scripts/gpu_secondary.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try:
        r=json.loads(l); print(r['config'][:12], r['dtype'], round(r['leapfrog_steps_per_s']/1e6,1))
    except Exception: pass"
done; done

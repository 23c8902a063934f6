"""One sampling launch of k_advance on the funnel (for ncu): 8192 chains, D=100, fixed step size."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import inplacedhmc_jl_b200 as bn
dt = bn.F64 if os.environ.get("DT", "f64") == "f64" else bn.F32
e = bn.Engine(8192, 100, dtype=dt, seed=11)
e.model_funnel(); e.set_positions(None); e.set_stepsize(0.15)
torch.cuda.synchronize(); t = time.perf_counter()
e.sample_device_only(int(os.environ.get("DRAWS", "20")))
torch.cuda.synchronize(); dt_s = time.perf_counter() - t
c = e.counters()
print(f"leapfrogs {c['leapfrogs']} in {dt_s*1e3:.1f} ms -> {c['leapfrogs']/dt_s/1e6:.1f} M/s")

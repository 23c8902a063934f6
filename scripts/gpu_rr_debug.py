import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import inplacedhmc_jl_b200 as bn
from bench import synth
N, D, C = int(os.environ.get("NROWS", 200000)), int(os.environ.get("DIM", 256)), 4096
bits, y, beta = synth(N, D)
e = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR); e.model_logistic(bits, y, 1.0)
rng = np.random.default_rng(1)
e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 2e-3)
b = e.get_state()[0].mean(0)
for _ in range(3):
    e.set_positions(np.tile(b, (C, 1))); b = b + e.get_state()[1][0] / (0.2 * N)
p = rng.normal(size=(C, D))
for mode in ("0", "1", "0", "1"):
    os.environ["BNUTS_TC_RREF"] = mode
    e.logistic_set_reference(b)
    e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 2e-3)
    e.leapfrog(p, 1e-3, 3)
    for n in (1, 5, 20):
        torch.cuda.synchronize(); t = time.perf_counter(); e.leapfrog(p, 1e-3, n); torch.cuda.synchronize()
        print(f"RREF={mode} nsteps={n:2d}: {(time.perf_counter() - t) * 1e3:8.2f} ms per call", flush=True)
    t = time.perf_counter(); e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 2e-3); torch.cuda.synchronize()
    print(f"RREF={mode} set_positions: {(time.perf_counter() - t) * 1e3:8.2f} ms", flush=True)

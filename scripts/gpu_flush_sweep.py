"""Accuracy and time of the tensor-path gradient vs accumulator flush period (run under gpurun)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import inplacedhmc_jl_b200 as bn
from bench import synth
N, D = 1_000_000, 100
bits, y, beta = synth(N, D)
X = (bits.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
rng = np.random.default_rng(4)
Cs = 128
q = np.zeros((Cs, D))
q[:64] = (beta + rng.normal(size=(64, D)) * 0.002).astype(np.float32)     # posterior bulk
q[64:] = (beta + rng.normal(size=(64, D)) * 0.05).astype(np.float32)      # far
eta = X @ q[[0, 1, 64, 65]].T
gref = ((y[:, None] - 1 / (1 + np.exp(-eta))).T @ X) - q[[0, 1, 64, 65]]
lref = (y[:, None] * eta - np.logaddexp(0, eta)).sum(0) - 0.5 * (q[[0, 1, 64, 65]] ** 2).sum(1)
print("|g| bulk %.1f %.1f far %.1f %.1f" % tuple(np.linalg.norm(gref, axis=1)))
for fl in os.environ.get("FLUSH", "4,16,32,64,128,0").split(","):
    os.environ["BNUTS_TC_FLUSH"] = fl
    for ns in os.environ.get("NSPLITS", "148,9").split(","):
        os.environ["BNUTS_TC_NSPLIT"] = ns
        e = bn.Engine(Cs, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR); e.model_logistic(bits, y, 1.0); e.set_positions(q)
        _, g, l = e.get_state()
        err = np.linalg.norm(g[[0, 1, 64, 65]] - gref, axis=1)
        print(f"flush {fl:>3s} nsplit {ns:>3s}: abs err bulk {err[0]:.4f} {err[1]:.4f} (rel {err[0]/np.linalg.norm(gref[0]):.2e}) "
              f"far {err[2]:.4f} {err[3]:.4f} (rel {err[2]/np.linalg.norm(gref[2]):.2e}); l abs err {np.abs(l[[0,1,64,65]]-lref).max():.4f}", flush=True)
        e.close()
    os.environ.pop("BNUTS_TC_NSPLIT")
    C = 4096
    e = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR); e.model_logistic(bits, y, 1.0)
    e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 2e-3)
    p = rng.normal(size=(C, D)); e.leapfrog(p, 1e-3, 3); e.profile(True); e.leapfrog(p, 1e-3, 20); ms, n = e.profile(False)
    print(f"flush {fl:>3s}: C=4096 grad kernel {ms/n*1e3:8.1f} us/launch ({4.0*N*D*C/(ms/n*1e-3)/1e12:6.1f} TFLOP/s alg)", flush=True)
    e.close()

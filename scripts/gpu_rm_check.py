"""Remainder mode of the tensor path (csrc/logistic_rm.cu) against numpy Float64: gradient and log density at points
near the reference (Taylor form), far from it (closed forms), mixed in one launch, small and ragged launches.
  gpurun -- 'python scripts/gpu_rm_check.py'          (NROWS, DIM override the shape)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth  # noqa: E402
import inplacedhmc_jl_b200 as bn  # noqa: E402

os.environ.setdefault("BNUTS_TC_RMODE", "2")   # the engine takes the mode by itself only for N >= 3000 D
N, D = int(os.environ.get("NROWS", 200_000)), int(os.environ.get("DIM", 100))
bits, y, beta = synth(N, D)
X = (bits.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
Xs = X * (2 * y - 1)[:, None]
b = beta.copy()
for _ in range(8):
    s = 1 / (1 + np.exp(-(X @ b)))
    H = (X * (s * (1 - s))[:, None]).T @ X + np.eye(D)
    b = b + np.linalg.solve(H, X.T @ (y - s) - b)
sd = 1.0 / np.sqrt(np.diag(H))


def ref(q):
    eta = Xs @ q.T                                           # [N][C]
    l = (np.minimum(eta, 0) - np.log1p(np.exp(-np.abs(eta)))).sum(0) - 0.5 * (q * q).sum(1)
    g = (Xs.T @ (1 / (1 + np.exp(eta)))).T - q
    return g, l


rng = np.random.default_rng(5)
worst = 0.0
for name, C, scale in [("near 1 sd, 256 chains", 256, 1.0), ("near 3 sd, 300 chains (ragged tile)", 300, 3.0),
                       ("40 chains (64-chain tiles)", 40, 1.0), ("7 chains", 7, 2.0), ("far 60 sd", 128, 60.0),
                       ("very far 1000 sd", 64, 1000.0), ("mixed near / far", 256, None), ("130 chains at the mode", 130, 0.0)]:
    if scale is None:
        sc = np.where(rng.uniform(size=C) < 0.1, 80.0, 1.0)[:, None]
    else:
        sc = scale
    q = (b[None, :] + rng.normal(size=(C, D)) * sd[None, :] * sc).astype(np.float32).astype(np.float64)
    g0, l0 = ref(q)
    e = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR)
    e.model_logistic(bits, y, 1.0)
    e.set_positions(q); _, g3, l3 = e.get_state()          # exact three-term path
    e.logistic_set_reference(b)
    t0 = time.perf_counter(); e.set_positions(q); dt = time.perf_counter() - t0
    _, g, l = e.get_state()
    gn = np.linalg.norm(g0, axis=1)
    eg = np.linalg.norm(g - g0, axis=1) / gn
    eg3 = np.linalg.norm(g3 - g0, axis=1) / gn
    el = np.abs(l - l0); el3 = np.abs(l3 - l0)
    print(f"{name:40s} grad rel err max {eg.max():.2e} median {np.median(eg):.2e} (three-term path {eg3.max():.2e}); "
          f"log density abs err max {el.max():.2e} (three-term {el3.max():.2e}); |grad| median {np.median(gn):.3g}", flush=True)
    worst = max(worst, eg.max())
    e.close()
print("worst gradient error", worst)

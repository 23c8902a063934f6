"""k_logistic_tc256 (128 < D <= 256) far from its reference point: gradient and log density against numpy Float64 at
positions of growing norm (the optimiser of config 5 starts at U[-2,2]^256, |eta| ~ 20)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import inplacedhmc_jl_b200 as bn
N, D, C = int(os.environ.get("NROWS", 60_000)), int(os.environ.get("DIM", 256)), 8
Xb, y, beta = bn.synth_logistic_rows(5, 0, N, D)
X = (Xb.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
Xs = X * (2 * y - 1)[:, None]
rng = np.random.default_rng(2)
e = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR)
e.model_logistic_synthetic(5, 0, N, 1.0)
for k in (0.05, 0.3, 1.0, 2.0, 5.0, 20.0):
    q = (rng.uniform(-1, 1, size=(C, D)) * k).astype(np.float32).astype(np.float64)
    eta = Xs @ q.T
    l0 = (np.minimum(eta, 0) - np.log1p(np.exp(-np.abs(eta)))).sum(0) - 0.5 * (q * q).sum(1)
    g0 = (Xs.T @ (1 / (1 + np.exp(eta)))).T - q
    e.set_positions(q)
    _, g, l = e.get_state()
    print("scale %5.2f  |eta| rms %7.2f max %8.2f  grad rel err %.2e  log density rel err %.2e  (l %.6g vs %.6g)" % (
        k, np.sqrt((eta ** 2).mean()), np.abs(eta).max(), np.max(np.linalg.norm(g - g0, axis=1) / np.linalg.norm(g0, axis=1)),
        np.max(np.abs(l - l0) / np.abs(l0)), l[0], l0[0]), flush=True)

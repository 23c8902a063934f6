"""Diagnose ragged trees on config 3 (run under gpurun)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import inplacedhmc_jl_b200 as bn
from bench import synth

N, D, C = 1_000_000, 100, int(os.environ.get("C", "1024"))
bits, y, beta = synth(N, D)
e = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR); e.model_logistic(bits, y, 1.0)
rng = np.random.default_rng(100)
e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 2e-3)
t = time.time(); e.find_initial_stepsize(); print("search %.2fs" % (time.time() - t), "eps0 quantiles", np.quantile(e.get_stepsize(), [0, .01, .5, .99, 1]))
t = time.time(); ch, st, ep = e.warmup_stage(100, bn.METRIC_NONE); print("adapt %.2fs" % (time.time() - t))
eps = e.get_stepsize()
print("eps quantiles", np.quantile(eps, [0, .001, .01, .1, .5, .9, .99, 1]))
print("warmup depth hist", np.bincount(st["depth"].ravel(), minlength=11))
print("warmup steps by transition index (max over chains):", st["steps"].max(0)[:20], "...", st["steps"].max(0)[-10:])
print("eps history of the chain with the largest single tree:")
cw = np.unravel_index(st["steps"].argmax(), st["steps"].shape)[0]
print(" chain", cw, "eps hist", ep[cw, ::10], "steps", st["steps"][cw, ::10], "acc", st["acceptance_rate"][cw, ::10])
t = time.time(); ch, st, sel = e.sample(8, want_index=True); dt = time.time() - t
print("sample(8): %.2fs, leapfrogs %d -> %.3e /s" % (dt, st["steps"].sum(), st["steps"].sum() / dt))
print("depth hist", np.bincount(st["depth"].ravel(), minlength=11))
tot = st["steps"].sum(1)
print("per-chain total steps quantiles", np.quantile(tot, [0, .5, .9, .99, .999, 1]))
worst = np.argsort(tot)[-5:]
for c in worst:
    print(" chain", c, "eps %.3e" % eps[c], "steps", st["steps"][c], "depth", st["depth"][c], "acc", np.round(st["acceptance_rate"][c], 2),
          "term", list(zip(st["term_left"][c], st["term_right"][c])))
print("typical chain eps %.3e" % np.median(eps), "acc mean", st["acceptance_rate"].mean())
print(e.counters())

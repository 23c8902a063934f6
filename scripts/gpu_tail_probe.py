#!/usr/bin/env python
"""Per-chain tree-length record of the c3 run (what bounds active_row_fraction): the bench set-up, then `calls`
calls of `T` transitions with statistics out; saves steps / depth per chain and transition plus the adapted step
sizes, so scheduling policies (barrier per call, one long call, run-ahead) can be replayed offline.
  gpurun -- 'python scripts/gpu_tail_probe.py --out gpurun_out/tail_probe.npz'"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=4096)
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=100)
ap.add_argument("--T", type=int, default=64)
ap.add_argument("--calls", type=int, default=12)
ap.add_argument("--adapt", type=int, default=40)
ap.add_argument("--full-warmup", action="store_true", help="75|25,50,100,200,400|50 as default_warmup_stages")
ap.add_argument("--out", default="gpurun_out/tail_probe.npz")
a = ap.parse_args()

import inplacedhmc_jl_b200 as bn  # noqa: E402

bits, y, beta = synth(a.rows, a.dim)
e = bn.Engine(a.chains, a.dim, dtype=bn.F32, seed=20261018, gradient_path=bn.GRAD_TENSOR)
e.model_logistic(bits, y, 1.0)
e.set_positions(None)
e.find_local_optimum(1e-4, 50)
e.logistic_set_reference(e.get_state()[0].mean(axis=0))
e.find_initial_stepsize()
stages = ((75, 0), (25, 1), (50, 1), (100, 1), (200, 1), (400, 1), (50, 0)) if a.full_warmup else \
    ((a.adapt, 0), (25, 1), (50, 1), (100, 1), (a.adapt, 0))
t0 = time.perf_counter()
for n, mk in stages:
    e.warmup_stage(n, mk, keep=False)
print("warmup %.1f s" % (time.perf_counter() - t0), flush=True)
steps = np.zeros((a.chains, a.calls * a.T), dtype=np.int16)
depth = np.zeros((a.chains, a.calls * a.T), dtype=np.int8)
lock = []
for k in range(a.calls):
    c0 = e.counters(); t0 = time.perf_counter()
    ch, st = e.sample(a.T)
    dt = time.perf_counter() - t0; c1 = e.counters()
    steps[:, k * a.T:(k + 1) * a.T] = st["steps"]; depth[:, k * a.T:(k + 1) * a.T] = st["depth"]
    lock.append((c1["lockstep_steps"] - c0["lockstep_steps"], c1["leapfrogs"] - c0["leapfrogs"], dt))
    print(k, lock[-1], flush=True)
np.savez_compressed(a.out, steps=steps, depth=depth, eps=e.get_stepsize(), minv=e.get_metric_diag().astype(np.float32),
                    lock=np.array(lock))
s = steps.astype(np.int64)
per_call = s.reshape(a.chains, a.calls, a.T).sum(axis=2)
print("mean steps/transition %.2f; barrier-per-call active fraction %.3f; one long call %.3f" % (
    s.mean(), per_call.mean() / per_call.max(axis=0).mean(), s.sum(axis=1).mean() / s.sum(axis=1).max()))
pc = s.mean(axis=1)
print("per-chain mean steps: min %.1f median %.1f p99 %.1f max %.1f" % (pc.min(), np.median(pc), np.percentile(pc, 99), pc.max()))

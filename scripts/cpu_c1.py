"""BASELINE config 1 — the reference's own CPU-runnable case: 100-dim iid standard normal, diagonal GaussianKineticEnergy,
NUTS max depth 10, default_warmup_stages (75 + 25..400 doubling + 50 steps, src/warmup.jl:361-372) + 1000 draws, ONE
chain — on the CPU oracle (the restatement of the reference algorithm; Julia is not in the image).  Prints one JSON line:
leapfrog steps/s and min-ESS/s of the sampling phase, posterior moments against N(0, I) within Monte-Carlo error."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import inplacedhmc_jl_b200 as bn
from conftest import build_oracle

lib = bn.load_library(build_oracle())
D, N = 100, 1000
stages = bn.default_warmup_stages(local_optimization=None, terminating_steps=150)   # 75 + 775 + 150 = 1000 warmup; q0 ~ U[-2,2] (src/warmup.jl:73); FindLocalOptimum is trivial here
t0 = time.perf_counter()
r = bn.mcmc_keep_warmup(bn.IIDNormal(D), N, warmup_stages=stages, nchains=1, lib=lib, seed=1)
t_all = time.perf_counter() - t0
e = bn.Engine(1, D, lib=lib, seed=1); bn.IIDNormal(D).attach(e)
dt = 1e9
for _ in range(7):                                               # the sampling phase alone, best of 7 identical re-runs
    e.restore(r["final_warmup_state"])
    t0 = time.perf_counter(); ch, st = e.sample(N); dt = min(dt, time.perf_counter() - t0)
    assert ch.tobytes() == r["inference"][0].tobytes()           # the timed re-run is the same chain (resume is exact)
warm = sum(w["stage"].N for w in r["warmup"] if hasattr(w["stage"], "N"))
ess = bn.diagnostics.min_ess(ch)
m, v = ch[0].mean(0), ch[0].var(0, ddof=1)
print(json.dumps({"config": "c1: iid N(0, I_100), diagonal metric, max_depth 10, %d warmup + %d draws, 1 chain, CPU oracle (1 core)" % (warm, N),
                  "leapfrog_steps": int(st["steps"].sum()), "sampling_s": dt, "leapfrog_steps_per_s": float(st["steps"].sum() / dt),
                  "min_ess": float(ess), "min_ess_per_s": float(ess / dt), "whole_run_s": t_all,
                  "mean_tree_depth": float(st["depth"].mean()), "mean_acceptance": float(st["acceptance_rate"].mean()),
                  "stepsize": float(r["final_warmup_state"]["ϵ"][0]),
                  "max_abs_mean": float(np.abs(m).max()), "var_range": [float(v.min()), float(v.max())],
                  "divergences": int((st["term_left"] == st["term_right"]).sum())}))

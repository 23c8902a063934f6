"""Throughput of the other BASELINE configs on one B200 (run under gpurun): c4 funnel (8192 chains),
iid normal (4096 chains), c2 dense Gaussian D=1000 (4096 chains, deterministic CUDA-core gradient, identity/diag metric)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import inplacedhmc_jl_b200 as bn
from conftest import make_gaussian
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

def run(name, C, D, dtype, setup, n_draws, stages=((75, 0), (25, 1), (50, 1), (100, 1), (50, 0)), delta=0.8, eps=None, flops_per_leapfrog=0, path=bn.GRAD_DETERMINISTIC):
    e = bn.Engine(C, D, dtype=dtype, seed=11, gradient_path=path, lib=bn.load_library(os.environ['BNUTS_LIB']) if os.environ.get('BNUTS_LIB') else None)
    setup(e); e.set_positions(None)
    t = time.perf_counter()
    if eps is None:
        e.find_initial_stepsize()
        for n, mk in stages:
            e.warmup_stage(n, mk, delta=delta, keep=False)
    else:
        e.set_stepsize(eps)
    tw = time.perf_counter() - t
    c0 = e.counters(); torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True); ev0.record()
    e.sample_device_only(n_draws)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1); c1 = e.counters()
    leap = c1["leapfrogs"] - c0["leapfrogs"]
    sz = 8 if dtype == bn.F64 else 4
    rec = {"config": name, "chains": C, "dim": D, "dtype": "f64" if dtype == bn.F64 else "f32", "draws": n_draws, "warmup_s": round(tw, 2),
           "sample_ms": round(ms, 2), "leapfrog_steps_per_s": leap / (ms * 1e-3), "mean_leapfrogs_per_transition": leap / (n_draws * C),
           "divergences": c1["divergences"] - c0["divergences"], "lockstep_steps": c1["lockstep_steps"] - c0["lockstep_steps"],
           "hbm_alg_GBs": leap * 13 * D * sz / (ms * 1e-3) / 1e9, "hbm_peak_GBs": PEAK}
    rec["hbm_frac"] = rec["hbm_alg_GBs"] / PEAK
    if flops_per_leapfrog:
        rec["alg_TFLOPs"] = leap * flops_per_leapfrog / (ms * 1e-3) / 1e12
    print(json.dumps(rec), flush=True)
    e.close()

if __name__ == "__main__":
    which = os.environ.get("WHICH", "funnel,iid,gauss").split(",")
    if "funnel" in which:
        for dt in (bn.F64, bn.F32):
            run("c4 funnel D=100", 8192, 100, dt, lambda e: e.model_funnel(), 100, delta=0.9)
    if "iid" in which:
        for dt in (bn.F64, bn.F32):
            run("c1-shape iid normal D=100 (batched)", 4096, 100, dt, lambda e: e.model_iid_normal(), 200)
    if "gauss" in which:
        P, S = make_gaussian(1000)
        run("c2 dense Gaussian D=1000, diag metric, CUDA-core deterministic gradient", 4096, 1000, bn.F32, lambda e: e.model_gaussian(P), 5,
            stages=((20, 0),), flops_per_leapfrog=2 * 1000 * 1000)

        def dense(e):
            e.model_gaussian(P); e.set_metric_dense(S)     # c2: dense-metric GaussianKE, M⁻¹ = Σ injected
        run("c2 dense Gaussian D=1000, shared dense metric M^-1 = Sigma (whitened: one GEMM per leapfrog), CUDA-core gradient", 4096, 1000, bn.F32, dense, 20,
            stages=((75, 0), (100, 0)), flops_per_leapfrog=2 * 1000 * 1000)
        run("c2 dense Gaussian D=1000, shared dense metric M^-1 = Sigma, tcgen05 gradient (3x3-term bf16 split)", 4096, 1000, bn.F32, dense, 20,
            stages=((75, 0), (100, 0)), flops_per_leapfrog=2 * 1000 * 1000, path=bn.GRAD_TENSOR)
        run("c2 dense Gaussian D=1000, diag metric, tcgen05 gradient", 4096, 1000, bn.F32, lambda e: e.model_gaussian(P), 5,
            stages=((20, 0),), flops_per_leapfrog=2 * 1000 * 1000, path=bn.GRAD_TENSOR)

"""Gradient-kernel time vs active rows (run under gpurun).  BNUTS_LIB selects a library variant,
REF=1 sets a reference point (two-term position operand)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import inplacedhmc_jl_b200 as bn
from bench import synth, ClockSampler
EPS = float(os.environ.get("EPS", 1e-5))   # small: the chains stay where they were put (the remainder mode has a near and a far regime)
N, D = int(os.environ.get("NROWS", 1_000_000)), int(os.environ.get("DIM", 100))
bits, y, beta = synth(N, D)
lib = bn.load_library(os.environ["BNUTS_LIB"]) if os.environ.get("BNUTS_LIB") else None
for C in [int(c) for c in os.environ.get("CS", "4096,2048,1024,512,128,16").split(",")]:
    e = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR, lib=lib); e.model_logistic(bits, y, 1.0)
    rng = np.random.default_rng(1)
    e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 2e-3)
    if os.environ.get("REF") == "1":
        if os.environ.get("REF_FILE") and os.path.exists(os.environ["REF_FILE"]):
            b = np.load(os.environ["REF_FILE"])          # ablated builds: the point found by the default library
        else:
            q, g, l = e.get_state()
            b = q.mean(0)
            for _ in range(3):   # crude ascent towards the mode using the engine's own gradient (H ~ N/4 x'x ~ 0.2 N I)
                e.set_positions(np.tile(b, (C, 1))); b = b + e.get_state()[1][0] / (0.2 * N)
            if os.environ.get("REF_FILE"):
                np.save(os.environ["REF_FILE"], b)
        e.logistic_set_reference(b)
        e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 2e-3)
    p = rng.normal(size=(C, D))
    e.leapfrog(p, EPS, 3)
    e.profile(True)
    NL = int(os.environ.get("NLEAP", 20))
    clk = ClockSampler(0); clk.start()
    torch.cuda.synchronize(); t = time.perf_counter(); e.leapfrog(p, EPS, NL); torch.cuda.synchronize(); dt = time.perf_counter() - t
    ck = clk.stop()
    ms, n = e.profile(False)
    fl = 4.0 * N * D * C
    print(f"{os.path.basename(os.environ.get('BNUTS_LIB', 'default'))} ref={os.environ.get('REF', '0')} C={C:5d}: grad kernel {ms/n*1e3:8.1f} us/launch ({fl/(ms/n*1e-3)/1e12:7.1f} TFLOP/s alg), lockstep step {dt/NL*1e3:7.3f} ms wall, sm {ck['sm_mhz']} MHz {ck['reasons']}", flush=True)
    e.close()

"""First-contact diagnostics for the tcgen05 logistic kernel (run under gpurun)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import inplacedhmc_jl_b200 as bn
from conftest import make_logistic, ORACLE_SO

def rel(a, b): return np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-30)

def grad_check(N, D, C, structured=False):
    X, y, beta = make_logistic(N, D)
    if structured:
        X = np.zeros((N, D)); X[np.arange(N), np.arange(N) % D] = 1.0
    rng = np.random.default_rng(1)
    q = beta[None, :] + rng.normal(size=(C, D)) * 0.3
    ref = bn.Engine(C, D, dtype=bn.F64, lib=ORACLE_SO); ref.model_logistic(X, y, 1.0); ref.set_positions(q)
    tc = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR); tc.model_logistic(X, y, 1.0)
    tc.set_positions(q)
    _, g0, l0 = ref.get_state(); _, g1, l1 = tc.get_state()
    print(f"N={N} D={D} C={C} structured={structured}: grad rel max {rel(g1, g0).max():.3e} median {np.median(rel(g1,g0)):.3e}; "
          f"l rel max {np.max(np.abs(l1-l0)/np.abs(l0)):.3e}", flush=True)
    if rel(g1, g0).max() > 1e-3:
        print(" g0[0,:8]", g0[0, :8]); print(" g1[0,:8]", g1[0, :8]); print(" l0[:4]", l0[:4], "l1[:4]", l1[:4])

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for args in [(128, 64, 128, True), (128, 64, 128, False), (256, 100, 128, True), (2000, 100, 200, False), (50000, 100, 256, False)]:
        try:
            grad_check(*args)
        except Exception as ex:
            print("FAILED", args, repr(ex), flush=True)
    # accumulator flush experiment at a larger N
    for fl in ("0", "4", "8", "32"):
        os.environ["BNUTS_TC_FLUSH"] = fl
        print("flush_every", fl, end=": ")
        try:
            grad_check(400000, 100, 128)
        except Exception as ex:
            print("FAILED", repr(ex), flush=True)
    os.environ.pop("BNUTS_TC_FLUSH", None)
    # timing at BASELINE config 3 size
    N, D, C = 1_000_000, 100, 4096
    X, y, beta = make_logistic(N, D)
    q = np.tile(beta, (C, 1))
    tc = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR); tc.model_logistic(X, y, 1.0)
    tc.set_positions(q); tc.set_stepsize(0.005)
    p = np.random.default_rng(0).normal(size=(C, D))
    for n in (1, 8, 32):
        torch.cuda.synchronize(); t = time.time(); tc.leapfrog(p, 0.005, n); torch.cuda.synchronize()
        dt = time.time() - t
        print(f"c3 bare leapfrog x{n}: {dt*1e3:.1f} ms total, {dt/n*1e3:.2f} ms/step, {C*n/dt:.3e} chain-leapfrogs/s", flush=True)
    t = time.time(); ch, st = tc.sample(2); dt = time.time() - t
    print(f"c3 sample(2): {dt:.2f}s, steps {st['steps'].sum()}, {st['steps'].sum()/dt:.3e} leapfrogs/s, depth hist {np.bincount(st['depth'].ravel())}")

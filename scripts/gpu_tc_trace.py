"""Dump the clock64 trace of one CTA of k_logistic_tc (debug build with -DBNUTS_TC_TRACE; run under gpurun)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import inplacedhmc_jl_b200 as bn
from bench import synth
N, D, C = 1_000_000, 100, 4096
bits, y, beta = synth(N, D)
lib = bn.load_library(os.path.join(ROOT, "build", os.environ.get("TRACELIB", "libbnuts_trace.so")))
e = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR, lib=lib); e.model_logistic(bits, y, 1.0)
rng = np.random.default_rng(1)
q = beta[None, :] + rng.normal(size=(C, D)) * 2e-3
for _ in range(3):
    e.set_positions(q)
t = np.zeros(3 * 256 * 16, dtype=np.int64)
rc = lib.bnuts_debug_tc_trace(t.ctypes.data_as(ctypes.c_void_p)); assert rc == 0
t = t.reshape(3, 256, 16)
np.save(os.path.join(ROOT, "gpurun_out", "tc_trace.npy"), t)
t0 = t[1, 0, 0]
for i in range(40, 52):
    m, w = t[1, i] - t0, t[2, i] - t0
    print(f"blk {i}: G1 wait {m[0]}..{m[1]} mmas ..{m[7]} commit ..{m[2]} | G2 rwait {m[3]}..{m[4]} g {m[5]} mmas ..{m[8]} commit ..{m[6]} end {m[9]} | EW sfull {w[0]}..{w[1]} ld {w[3]}..{w[4]} compute ..{w[5]} st ..{w[6]}..{w[7]} arrive {w[2]}")

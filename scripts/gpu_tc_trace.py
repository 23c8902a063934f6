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
for i in range(40, 56):
    p, m, w = t[0, i] - t0, t[1, i] - t0, t[2, i] - t0
    print(f"  it {i}: xwait {m[0]}..{m[1]} rwait ..{m[4]} g {m[5]} | G1({i+2}) mmas {t[1,i+2,7]-t0} commit {t[1,i+2,2]-t0} | G2({i}) mmas {m[8]} commit {m[6]} loopend {m[9]}")
    print(f"  EW {i}: ldwait {w[3]}..{w[4]} compute ..{w[5]} prefetch+st issue ..{w[6]} stwait ..{w[7]} arrive {w[2]} | next sfull wait {t[2,i+1,0]-t0}..{t[2,i+1,1]-t0}")
    print(f"blk {i}: TMA wait {p[0]}..{p[1]} | G1: xfull wait {m[0]}..{m[1]} issued {m[2]} | G2: rfull wait {m[3]}..{m[4]} gwait {m[5]} issued {m[6]} | EW: sfull wait {w[0]}..{w[1]} arrive {w[2]}")

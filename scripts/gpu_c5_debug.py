"""c5 set-up path of bench.py on one GPU at a reduced row count: state of the chains after every phase."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import inplacedhmc_jl_b200 as bn
N = int(os.environ.get("NROWS", 2_000_000)); D = 256; C = int(os.environ.get("CHAINS", 4096))
e = bn.Engine(C, D, dtype=bn.F32, seed=20261018, gradient_path=bn.GRAD_TENSOR)
e.model_logistic_synthetic(5, 0, N, 1.0)
_, _, beta = bn.synth_logistic_rows(5, 0, 0, D)
def show(tag):
    q, g, l = e.get_state()
    st = e.chain_status()
    print(tag, "status counts", dict(zip(*np.unique(st, return_counts=True))), "|q - beta*| median %.3g max %.3g" % (np.median(np.linalg.norm(q - beta, axis=1)), np.linalg.norm(q - beta, axis=1).max()),
          "|grad| median %.3g max %.3g" % (np.median(np.linalg.norm(g, axis=1)), np.nanmax(np.linalg.norm(g, axis=1))), "l min %.6g max %.6g nonfinite %d" % (np.nanmin(l), np.nanmax(l), (~np.isfinite(l)).sum()),
          "spread of chains about their mean: %.3g" % np.linalg.norm(q - q.mean(0), axis=1).max(), flush=True)
e.set_positions(None); show("init")
for r in range(3):
    try:
        e.find_local_optimum(1e-4, int(os.environ.get("OPT_ITERS", 30)))
    except bn.BnutsError as ex:
        print("optimum:", ex)
    show("after optimum %d" % r)
    e.logistic_set_reference(e.get_state()[0].mean(axis=0)); show("after set_reference %d" % r)
try:
    e.find_initial_stepsize()
except bn.BnutsError as ex:
    print("search:", ex)
show("after search")
eps = e.get_stepsize(); print("eps", np.percentile(eps, [0, 50, 100]))

"""Generate the golden fixtures under tests/golden/ (committed; rerun only on purpose).

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference (chriselrod/InplaceDHMC.jl) ships no tests, golden vectors or
fixtures (test/runtests.jl:4-6 is an empty testset) and cannot run here (no Julia; un-vendored
dependencies), so these vectors cannot come from the reference itself.  Provenance of each file:

  tree_decisions_iid.npz   the independent literal Python transcription of the reference recursion
                           (tests/test_tree_semantics.py::Ref, following src/tree.jl:321-444,
                           src/NUTS.jl:148-191) with injected momenta / directions, the reference's own
                           test hooks (src/NUTS.jl:251-258).  Independent of the oracle's tree code.
  logistic_numpy.npz       value and gradient of the logistic target from plain numpy Float64
                           (SURVEY.md §A.4), independent of every library in this repo.
  leapfrog_closed_form.npz leapfrog on N(0,I) with unit metric is a linear map: closed-form positions
                           after n steps (src/kinetic_energy.jl:144-161 restated as a 2x2 recurrence).
  protocol_<kind>_<dtype>.npz
                           the CPU oracle (oracle/bnuts_oracle.cpp) driven through the C ABI by the fixed
                           script tests/conftest.py::run_protocol (search, windowed warmup with metric
                           updates, draws, statistics, bare leapfrogs).  These pin the oracle itself
                           against drift and give the CUDA engine a bit-for-bit target that does not
                           need the oracle at test time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import inplacedhmc_jl_b200 as bn  # noqa: E402
from conftest import build_oracle, run_protocol, make_logistic  # noqa: E402
from test_tree_semantics import Ref  # noqa: E402

TREE_CASES = [(0.3, 5), (0.9, 4), (1.9, 6)]
TREE_SHAPE = dict(Cn=8, D=5, T=5, seed=4321)


def tree_inputs(k):
    rng = np.random.default_rng(100 + k)
    Cn, D, T = TREE_SHAPE["Cn"], TREE_SHAPE["D"], TREE_SHAPE["T"]
    q0 = rng.normal(size=(Cn, D))
    p = rng.normal(size=(T, Cn, D))
    dirs = rng.integers(0, 2 ** 32, size=(T, Cn), dtype=np.uint64).astype(np.uint32)
    return q0, p, dirs


def main():
    lib = bn.load_library(build_oracle())
    # ---- tree decisions from the literal transcription (teacher-forced on its own draws)
    out = {}
    for k, (eps, md) in enumerate(TREE_CASES):
        q0, p, dirs = tree_inputs(k)
        Cn, D, T, seed = TREE_SHAPE["Cn"], TREE_SHAPE["D"], TREE_SHAPE["T"], TREE_SHAPE["seed"]
        rec = np.zeros((Cn, T, 6), dtype=np.int64)        # term_left, term_right, depth, steps, selected index, -
        draws = np.zeros((Cn, T, D)); pis = np.zeros((Cn, T)); accs = np.zeros((Cn, T))
        for c in range(Cn):
            q = q0[c]
            for t in range(T):
                zeta, pi, acc, term, depth, steps = Ref(lib, seed, c, t, eps, md).sample(q, p[t, c], int(dirs[t, c]))
                rec[c, t, :5] = (term[0], term[1], depth, steps, zeta[1])
                draws[c, t] = zeta[0][0]; pis[c, t] = pi; accs[c, t] = acc
                q = zeta[0][0]
        out.update({f"rec{k}": rec, f"draws{k}": draws, f"pi{k}": pis, f"acc{k}": accs})
    np.savez_compressed(os.path.join(HERE, "tree_decisions_iid.npz"), **out)
    # ---- logistic target, numpy Float64
    N, D = 600, 24
    X, y, beta = make_logistic(N, D, seed=9)
    rng = np.random.default_rng(9)
    q = np.asarray(beta[None, :] + rng.normal(size=(16, D)) * 0.4, dtype=np.float32).astype(np.float64)
    eta = X @ q.T
    g = ((y[:, None] - 1 / (1 + np.exp(-eta))).T @ X) - 1.0 * q
    l = (y[:, None] * eta - np.logaddexp(0, eta)).sum(0) - 0.5 * (q ** 2).sum(1)
    np.savez_compressed(os.path.join(HERE, "logistic_numpy.npz"), X=X, y=y, q=q, grad=g, logdensity=l)
    # ---- closed-form leapfrog on N(0, I): (q, p) -> A^n (q, p), A = [[1 - e^2/2, e], [-e + e^3/4, 1 - e^2/2]]
    rng = np.random.default_rng(10)
    q0 = rng.normal(size=(4, 7)); p0 = rng.normal(size=(4, 7)); eps = 0.37; n = 9
    A = np.array([[1 - eps ** 2 / 2, eps], [-eps + eps ** 3 / 4, 1 - eps ** 2 / 2]])
    An = np.linalg.matrix_power(A, n)
    np.savez_compressed(os.path.join(HERE, "leapfrog_closed_form.npz"), q0=q0, p0=p0, eps=eps, n=n,
                        q=An[0, 0] * q0 + An[0, 1] * p0, p=An[1, 0] * q0 + An[1, 1] * p0)
    # ---- oracle protocol outputs
    for kind in ("iid", "funnel", "gauss", "logit"):
        for dtype, nm in ((bn.F64, "f64"), (bn.F32, "f32")):
            arrs = run_protocol(bn, lib, kind, dtype)
            np.savez_compressed(os.path.join(HERE, f"protocol_{kind}_{nm}.npz"), **{f"a{i:02d}": np.asarray(a) for i, a in enumerate(arrs)})
    synth_fixture(lib)
    print("golden fixtures written to", HERE)


SYNTH = dict(seed=5, D=10, blocks=((0, 6), (4294967294, 4), (12500000 * 7 + 3, 3)))   # row ranges incl. one across 2^32


def synth_fixture(lib):
    """Frozen rows of the synthetic design matrix (include/bnuts.h, bnuts_model_logistic_synthetic) from the oracle's
    host generator: pins the (seed, row, column) -> value map, including row indices beyond 32 bits."""
    out = {}
    for k, (r0, n) in enumerate(SYNTH["blocks"]):
        X, y, beta = bn.synth_logistic_rows(SYNTH["seed"], r0, n, SYNTH["D"], lib=lib)
        out.update({f"X{k}": X, f"y{k}": y})
    out["beta"] = beta
    np.savez_compressed(os.path.join(HERE, "synth_rows.npz"), **out)


if __name__ == "__main__":
    main()

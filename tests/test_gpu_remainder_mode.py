"""Remainder mode of the tensor-core logistic path (csrc/logistic_rm.cu): the model expanded about the reference point row
by row, the D x D linear / quadratic part exact in the consumer, only the Taylor remainder through the tensor cores.
Gradient and log density against numpy Float64 inside the Taylor radius, outside it (closed forms, two bf16 terms of the
remainder), mixed in one launch, for 128- and 64-chain tiles, ragged tiles and a handful of chains."""
import numpy as np
import pytest

from conftest import make_logistic

pytestmark = pytest.mark.gpu


def _setup(N, D):
    X, y, beta = make_logistic(N, D)
    Xs = X * (2 * y - 1)[:, None]
    b = beta.copy()
    for _ in range(8):
        s = 1 / (1 + np.exp(-(X @ b)))
        H = (X * (s * (1 - s))[:, None]).T @ X + np.eye(D)
        b = b + np.linalg.solve(H, X.T @ (y - s) - b)
    return X, y, Xs, b, 1.0 / np.sqrt(np.diag(H))


def _ref(Xs, q):
    eta = Xs @ q.T
    l = (np.minimum(eta, 0) - np.log1p(np.exp(-np.abs(eta)))).sum(0) - 0.5 * (q * q).sum(1)
    return (Xs.T @ (1 / (1 + np.exp(eta)))).T - q, l


@pytest.fixture(scope="module")
def problem():
    return _setup(200_000, 100)


@pytest.mark.parametrize("name,C,scale,gtol,ltol", [
    ("one posterior sd", 256, 1.0, 2e-6, 2e-3),
    ("three sd, ragged tile", 300, 3.0, 2e-6, 2e-3),
    ("64-chain tiles", 40, 1.0, 2e-6, 2e-3),
    ("a handful of chains", 7, 2.0, 2e-6, 2e-3),
    ("far: 60 sd (closed forms)", 128, 60.0, 1e-5, None),
])
def test_remainder_mode_matches_float64(bn, cuda_lib, monkeypatch, problem, name, C, scale, gtol, ltol):
    X, y, Xs, b, sd = problem
    D = X.shape[1]
    monkeypatch.setenv("BNUTS_TC_RMODE", "2")   # taken by itself only for N >= 3000 D
    rng = np.random.default_rng(5)
    q = (b[None, :] + rng.normal(size=(C, D)) * sd[None, :] * scale).astype(np.float32).astype(np.float64)
    g0, l0 = _ref(Xs, q)
    e = bn.Engine(C, D, dtype=bn.F32, lib=cuda_lib, gradient_path=bn.GRAD_TENSOR)
    e.model_logistic(X, y, 1.0)
    e.logistic_set_reference(b)
    e.set_positions(q)
    _, g, l = e.get_state()
    err = np.linalg.norm(g - g0, axis=1) / np.linalg.norm(g0, axis=1)
    assert err.max() < gtol, (name, err.max())
    if ltol is not None:
        assert np.abs(l - l0).max() < ltol, (name, np.abs(l - l0).max())
    else:
        assert np.max(np.abs(l - l0) / np.abs(l0)) < 1e-5, np.max(np.abs(l - l0) / np.abs(l0))   # Float32 sums of O(1) terms far from the reference
    # per-leapfrog parity with the mode in force: positions and gradients after three steps against Float64 leapfrogs on the host
    p = rng.normal(size=(C, D)).astype(np.float32).astype(np.float64) * 30.0
    eps, n = 2e-4, 3
    qa, pa, ga = q.copy(), p.copy(), g0.copy()
    for _ in range(n):
        pa = pa + 0.5 * eps * ga; qa = qa + eps * pa; ga, _ = _ref(Xs, qa); pa = pa + 0.5 * eps * ga
    qc, pc, gc, lc = e.leapfrog(p, eps, n)
    eq = np.max(np.abs(qc - qa) / (np.abs(qa) + 1e-3)); eg = np.max(np.linalg.norm(gc - ga, axis=1) / np.linalg.norm(ga, axis=1))
    assert eq < 1e-5 and eg < 5e-5, (name, eq, eg)
    e.close()


def test_remainder_mode_mixed_launch_and_slots(bn, cuda_lib, monkeypatch, problem):
    """Near and far chains in one launch: the near ones keep their accuracy next to far neighbours, and a position gives the
    same bits whatever its staging row (duplicates of five positions spread over the tiles, all tiles of one kind)."""
    X, y, Xs, b, sd = problem
    D, C = X.shape[1], 256
    monkeypatch.setenv("BNUTS_TC_RMODE", "2")
    rng = np.random.default_rng(6)
    sc = np.where(np.arange(C) % 9 == 4, 80.0, 1.0)[:, None]
    q = (b[None, :] + rng.normal(size=(C, D)) * sd[None, :] * sc).astype(np.float32).astype(np.float64)
    g0, l0 = _ref(Xs, q)
    e = bn.Engine(C, D, dtype=bn.F32, lib=cuda_lib, gradient_path=bn.GRAD_TENSOR)
    e.model_logistic(X, y, 1.0); e.logistic_set_reference(b); e.set_positions(q)
    _, g, l = e.get_state()
    err = np.linalg.norm(g - g0, axis=1) / np.linalg.norm(g0, axis=1)
    near = sc[:, 0] == 1.0
    assert err[near].max() < 2e-6 and err[~near].max() < 2e-5, (err[near].max(), err[~near].max())
    # slot invariance: five near positions repeated over all rows
    q2 = np.repeat(q[near][:5], C // 5 + 1, axis=0)[:C]
    e.set_positions(q2)
    _, g2, l2 = e.get_state()
    for k in range(5, C):
        j = k // (C // 5 + 1)
        first = j * (C // 5 + 1)
        assert g2[k].tobytes() == g2[first].tobytes() and l2[k] == l2[first]
    # leaving the mode restores the exact path
    e.logistic_set_reference(None)
    e.set_positions(q)
    _, g3, _ = e.get_state()
    assert np.max(np.linalg.norm(g3 - g0, axis=1) / np.linalg.norm(g0, axis=1)) < 2e-5
    e.close()


def test_remainder_mode_is_taken_for_tall_problems_only(bn, cuda_lib, monkeypatch):
    """N >= 3000 D switches the mode on by itself (posterior rms of x·(beta - beta0) below ~0.04); smaller problems keep the
    residual operands of k_logistic_tc.  Seen from outside through the accuracy of the log density (1e-4 vs 3e-2)."""
    monkeypatch.delenv("BNUTS_TC_RMODE", raising=False); monkeypatch.delenv("BNUTS_TC_RREF", raising=False)
    for N, D, expect in ((40_000, 10, True), (20_000, 10, False)):
        X, y, Xs, b, sd = _setup(N, D)
        C = 64
        q = (b[None, :] + np.random.default_rng(1).normal(size=(C, D)) * sd[None, :]).astype(np.float32).astype(np.float64)
        g0, l0 = _ref(Xs, q)
        e = bn.Engine(C, D, dtype=bn.F32, lib=cuda_lib, gradient_path=bn.GRAD_TENSOR)
        e.model_logistic(X, y, 1.0); e.logistic_set_reference(b); e.set_positions(q)
        _, g, l = e.get_state()
        assert np.max(np.linalg.norm(g - g0, axis=1) / np.linalg.norm(g0, axis=1)) < 1e-5
        tight = np.abs(l - l0).max() < 2e-4
        assert tight == expect or (not expect and tight), (N, np.abs(l - l0).max())   # a small N may be accurate either way; a tall one must be
        e.close()


@pytest.mark.parametrize("D,N", [(100, 40_000), (256, 60_000)])
def test_remainder_mode_with_sharded_rows_group_of_one(bn, cuda_lib, monkeypatch, D, N):
    """Rows "sharded" over a group of one (NCCL, world 1): the set-up constants go through the group sum, the folded
    remainder through the per-leapfrog exchange, and the consumer adds the linear part behind it — against the ordinary
    engine in the same mode and against numpy Float64 (the 2- and 8-GPU runs are scripts/gpu_c5_debug.py and bench.py --config c5)."""
    monkeypatch.setenv("BNUTS_TC_RMODE", "2")
    X, y, Xs, b, sd = _setup(N, D)
    C = 96
    rng = np.random.default_rng(9)
    q = (b[None, :] + rng.normal(size=(C, D)) * sd[None, :]).astype(np.float32).astype(np.float64)
    g0, l0 = _ref(Xs, q)
    outs = []
    for sharded in (False, True):
        e = bn.Engine(C, D, dtype=bn.F32, lib=cuda_lib, gradient_path=bn.GRAD_TENSOR, seed=4, max_depth=6)
        e.model_logistic(X, y, 1.0)
        if sharded:
            try:
                e.set_nccl(bn.nccl_unique_id(cuda_lib), 1, 0)
            except bn.BnutsError as ex:
                pytest.skip(f"NCCL not loadable here: {ex}")
        e.logistic_set_reference(b)
        e.set_positions(q)
        _, g, l = e.get_state()
        e.set_stepsize(1e-3)
        ch, st = e.sample(3)
        outs.append((g, l, ch, st))
        e.close()
    (ga, la, cha, sta), (gb, lb, chb, stb) = outs
    for g, l in ((ga, la), (gb, lb)):
        assert np.max(np.linalg.norm(g - g0, axis=1) / np.linalg.norm(g0, axis=1)) < 2e-6
        assert np.abs(l - l0).max() < 2e-3
    assert np.max(np.linalg.norm(gb - ga, axis=1) / np.linalg.norm(ga, axis=1)) < 3e-7 and np.abs(lb - la).max() < 1e-4
    assert (sta["steps"] == stb["steps"]).mean() > 0.9


@pytest.mark.parametrize("D", [61, 129, 200])
def test_remainder_mode_other_widths(bn, cuda_lib, monkeypatch, D):
    """Tile widths the headline shapes do not reach: D = 61 (one 64-column chunk), D = 129 and 200 (three chunks: the second
    128-feature half of the GEMM2 accumulator is only partly backed by data)."""
    monkeypatch.setenv("BNUTS_TC_RMODE", "2")
    X, y, Xs, b, sd = _setup(30_000, D)
    C = 70
    q = (b[None, :] + np.random.default_rng(11).normal(size=(C, D)) * sd[None, :] * 0.3).astype(np.float32).astype(np.float64)
    g0, l0 = _ref(Xs, q)
    e = bn.Engine(C, D, dtype=bn.F32, lib=cuda_lib, gradient_path=bn.GRAD_TENSOR)
    e.model_logistic(X, y, 1.0); e.logistic_set_reference(b); e.set_positions(q)
    _, g, l = e.get_state()
    err = np.linalg.norm(g - g0, axis=1) / np.linalg.norm(g0, axis=1)
    assert err.max() < 3e-6 and np.abs(l - l0).max() < 2e-3, (D, err.max(), np.abs(l - l0).max())
    e.close()

"""Shared dense metric (≙ the dense GaussianKineticEnergy constructor the reference keeps as a comment,
src/hamiltonian.jl:44; BASELINE config 2).  The oracle applies M⁻¹ directly (mat-vecs in the kinetic energy,
p♯, the drift and the momentum draw); the engine whitens (unit metric on q̃ = L⁻¹q, model P̃ = LᵀPL, coordinate
maps at the C ABI).  Agreement of the two checks the equivalence: per-leapfrog 1e-12 relative in fp64,
identical tree decisions, same statistics."""
import numpy as np
import pytest

from conftest import make_gaussian

F64, F32 = 0, 1


def _setup(bn, lib, C, D, dtype, Minv, P, q0, **kw):
    e = bn.Engine(C, D, dtype=dtype, max_depth=6, lib=lib, seed=11, **kw)
    e.model_gaussian(P)
    e.set_metric_dense(Minv)
    e.set_positions(q0)
    return e


def _problem(D, C, seed=3):
    P, S = make_gaussian(D, seed=seed)
    rng = np.random.default_rng(seed)
    B = rng.normal(size=(D, D)) * 0.15
    Minv = S + B @ B.T * 0.3 + 0.05 * np.eye(D)         # a good but imperfect preconditioner
    Minv = 0.5 * (Minv + Minv.T)
    q0 = rng.normal(size=(C, D))
    return P, Minv, q0


def _check(bn, ref_lib, lib, dtype, tol, **kw):
    D, C = 13, 9
    P, Minv, q0 = _problem(D, C)
    a = _setup(bn, ref_lib, C, D, dtype, Minv, P, q0)
    b = _setup(bn, lib, C, D, dtype, Minv, P, q0, **kw)
    np.testing.assert_allclose(b.get_metric_dense(), Minv, rtol=0, atol=0)
    # state in user coordinates
    for x, y in zip(a.get_state(), b.get_state()):
        np.testing.assert_allclose(y, x, rtol=tol, atol=tol)
    np.testing.assert_allclose(a.get_state()[1], -(q0 @ P), rtol=100 * tol, atol=100 * tol)
    # bare leapfrogs with a user-space momentum: q' = q + eps M⁻¹ p_m etc.
    rng = np.random.default_rng(5)
    p = rng.normal(size=(C, D))
    for eps, n in ((0.2, 1), (-0.1, 4)):
        for x, y in zip(a.leapfrog(p, eps, n), b.leapfrog(p, eps, n)):
            np.testing.assert_allclose(y, x, rtol=20 * tol, atol=20 * tol)
    # one-step closed form on the oracle side: q1 = q0 + eps M⁻¹ (p + eps/2 g0)
    q1 = a.leapfrog(p, 0.2, 1)[0]
    np.testing.assert_allclose(q1, q0 + 0.2 * (p + 0.1 * (-(q0 @ P))) @ Minv, rtol=100 * tol, atol=100 * tol)
    # transitions with injected momenta / directions (teacher-forced on the oracle's draws)
    T = 6
    same = total = 0
    q = q0
    for t in range(T):
        pin = rng.normal(size=(1, C, D))
        dirs = rng.integers(0, 2 ** 32, size=(1, C), dtype=np.uint64).astype(np.uint32)
        outs = []
        for e in (a, b):
            e.seed(11, t); e.set_positions(q); e.set_stepsize(0.35); e.inject(1, dirs, pin)
            outs.append(e.sample(1, want_index=True))
        (ca, sa, ia), (cb, sb, ib) = outs
        ok = (sa["depth"] == sb["depth"]) & (sa["steps"] == sb["steps"]) & (sa["term_left"] == sb["term_left"]) & \
             (sa["term_right"] == sb["term_right"]) & (ia == ib)
        same += int(ok.sum()); total += ok.size
        k = ok[:, 0]
        np.testing.assert_allclose(cb[k], ca[k], rtol=1e3 * tol, atol=1e3 * tol)
        np.testing.assert_allclose(sb["pi"][k], sa["pi"][k], rtol=1e3 * tol, atol=1e3 * tol)
        q = ca[:, 0]
    assert same >= (total if dtype == F64 else 0.9 * total), (same, total)
    # free-running: Philox momenta p = L⁻ᵀ z on both sides
    for e in (a, b):
        e.seed(11, 100); e.set_positions(q0); e.find_initial_stepsize()
    np.testing.assert_allclose(b.get_stepsize(), a.get_stepsize(), rtol=1e-6 if dtype == F64 else 0.3)
    if dtype == F64:
        ca, sa = a.sample(4); cb, sb = b.sample(4)
        np.testing.assert_array_equal(sa["steps"][:, 0], sb["steps"][:, 0])
        np.testing.assert_allclose(cb[:, 0], ca[:, 0], rtol=1e-8, atol=1e-8)
    with pytest.raises(bn.BnutsError):
        b.warmup_stage(20, 1)                           # diagonal adaptation on top of a dense metric: refused
    a.close(); b.close()


@pytest.mark.parametrize("dtype,tol", [(F64, 1e-12), (F32, 2e-5)])
def test_dense_metric_whitened_engine_matches_direct_oracle_cpu(bn, oracle_lib, hostemu_lib, dtype, tol):
    _check(bn, oracle_lib, hostemu_lib, dtype, tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(F64, 1e-12), (F32, 2e-5)])
def test_dense_metric_whitened_engine_matches_direct_oracle_cuda(bn, oracle_lib, cuda_lib, dtype, tol):
    _check(bn, oracle_lib, cuda_lib, dtype, tol, gradient_path=1)


def test_dense_metric_posterior_and_switching(bn, hostemu_lib):
    """Posterior moments in user coordinates; switching the metric keeps the positions; iid normal target."""
    D, C = 6, 48
    P, Minv, q0 = _problem(D, C, seed=8)
    S = np.linalg.inv(P)
    e = bn.Engine(C, D, max_depth=6, lib=hostemu_lib, seed=2)
    e.model_gaussian(P); e.set_positions(q0)
    e.set_metric_dense(S)                               # perfect preconditioner: P̃ = I
    np.testing.assert_allclose(e.get_state()[0], q0, atol=1e-12)
    e.find_initial_stepsize(); e.warmup_stage(60, 0, keep=False)
    ch, st = e.sample(150)
    x = ch.reshape(-1, D)
    assert np.max(np.abs(x.mean(0)) / np.sqrt(np.diag(S))) < 0.15
    assert np.max(np.abs(np.cov(x.T) - S)) / np.max(np.abs(S)) < 0.15
    assert st["depth"].mean() < 3.2                     # isotropic in whitened coordinates: short trees
    e.set_metric_dense(None)
    np.testing.assert_allclose(e.get_metric_dense(), np.eye(D))
    np.testing.assert_allclose(e.get_state()[0], ch[:, -1], atol=1e-10)
    # iid normal under a dense metric
    f = bn.Engine(4, D, lib=hostemu_lib); f.model_iid_normal(); f.set_metric_dense(Minv); f.set_positions(q0[:4])
    q, g, l = f.get_state()
    np.testing.assert_allclose(g, -q0[:4], atol=1e-12); np.testing.assert_allclose(l, -0.5 * (q0[:4] ** 2).sum(1), atol=1e-12)
    with pytest.raises(bn.BnutsError):
        f.model_funnel()

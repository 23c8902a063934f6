"""Known-answer tests that pin the CPU oracle (the reference ships no tests or golden
vectors — test/runtests.jl:4-6 is empty — so the oracle is validated analytically,
SURVEY.md §4)."""
import ctypes as C
import math

import numpy as np
import pytest

from conftest import make_gaussian

F64, F32 = 0, 1


def _probe(lib):
    lib.bnuts_oracle_exp.restype = C.c_double; lib.bnuts_oracle_exp.argtypes = [C.c_double]
    lib.bnuts_oracle_log.restype = C.c_double; lib.bnuts_oracle_log.argtypes = [C.c_double]
    lib.bnuts_oracle_log1p.restype = C.c_double; lib.bnuts_oracle_log1p.argtypes = [C.c_double]
    lib.bnuts_oracle_logaddexp.restype = C.c_double; lib.bnuts_oracle_logaddexp.argtypes = [C.c_double, C.c_double]
    lib.bnuts_oracle_expf.restype = C.c_float; lib.bnuts_oracle_expf.argtypes = [C.c_float]
    lib.bnuts_oracle_logf.restype = C.c_float; lib.bnuts_oracle_logf.argtypes = [C.c_float]
    lib.bnuts_oracle_normal.restype = C.c_double
    lib.bnuts_oracle_normal.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.bnuts_oracle_exponential.restype = C.c_double
    lib.bnuts_oracle_exponential.argtypes = [C.c_uint64] + [C.c_uint32] * 5
    lib.bnuts_oracle_directions.restype = C.c_uint32
    lib.bnuts_oracle_directions.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    return lib


def test_philox_known_answers(oracle_lib):
    """Random123 kat_vectors for philox4x32-10."""
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        oracle_lib.bnuts_oracle_philox(c, k, o)
        assert list(o) == want


def test_scalar_math_accuracy(oracle_lib):
    lib = _probe(oracle_lib)
    rng = np.random.default_rng(0)
    ulp = 2.220446049250313e-16
    for x in np.concatenate([rng.uniform(-700, 700, 4000), rng.uniform(-2, 2, 4000)]):
        assert abs(lib.bnuts_oracle_exp(x) / math.exp(x) - 1) < 3 * ulp
    for x in np.concatenate([np.exp(rng.uniform(-700, 700, 4000)), rng.uniform(0.5, 2, 4000)]):
        assert abs(lib.bnuts_oracle_log(x) - math.log(x)) <= 3 * ulp * max(abs(math.log(x)), 1e-300)
    for x in np.exp(rng.uniform(-40, 0, 4000)):
        assert abs(lib.bnuts_oracle_log1p(x) / math.log1p(x) - 1) < 4 * ulp
    for x in rng.uniform(-80, 80, 4000).astype(np.float32):
        assert abs(float(lib.bnuts_oracle_expf(float(x))) / math.exp(float(x)) - 1) < 3 * 1.19e-7
    assert lib.bnuts_oracle_exp(-1000.0) == 0.0 and lib.bnuts_oracle_exp(1000.0) == math.inf
    assert lib.bnuts_oracle_log(0.0) == -math.inf and math.isnan(lib.bnuts_oracle_log(-1.0))


def test_logaddexp_semantics(oracle_lib):
    """src/InplaceDHMC.jl:27-30: non-finite arguments return `x > y ? x : y`."""
    lib = _probe(oracle_lib)
    f = lib.bnuts_oracle_logaddexp
    inf, nan = math.inf, math.nan
    assert f(-inf, -inf) == -inf
    assert f(-inf, 1.5) == 1.5 and f(1.5, -inf) == 1.5
    assert f(nan, 1.0) == 1.0            # NaN > y is false -> returns y
    assert math.isnan(f(1.0, nan))       # x > NaN is false -> returns y = NaN
    assert f(inf, 0.0) == inf
    for x, y in [(0.0, 0.0), (-3.0, 2.0), (-700.0, -701.0), (5.0, -40.0)]:
        assert abs(f(x, y) - np.logaddexp(x, y)) < 1e-14 * max(1.0, abs(np.logaddexp(x, y)))


def test_rng_streams(oracle_lib):
    lib = _probe(oracle_lib)
    z = np.array([lib.bnuts_oracle_normal(1, c, 0, d) for c in range(100) for d in range(100)])
    assert abs(z.mean()) < 0.04 and abs(z.var() - 1) < 0.05
    assert abs(((z - z.mean()) ** 4).mean() / z.var() ** 2 - 3) < 0.2
    e = np.array([lib.bnuts_oracle_exponential(1, c, 3, 2, 1, 2) for c in range(5000)])
    assert abs(e.mean() - 1) < 0.05 and e.min() > 0
    bits = np.array([lib.bnuts_oracle_directions(1, c, 0) for c in range(2000)], dtype=np.uint64)
    assert abs(np.mean(bits & 1) - 0.5) < 0.05


def test_tree_stats_layout(bn):
    """≙ TreeStatisticsNUTS (src/NUTS.jl:229-242): π@0, acceptance_rate@8, termination@16/20, depth@24, steps@28."""
    dt = bn.TREE_STATS_DTYPE
    assert dt.itemsize == 32
    assert [dt.fields[n][1] for n in ("pi", "acceptance_rate", "term_left", "term_right", "depth", "steps")] == \
        [0, 8, 16, 20, 24, 28]


def test_leapfrog_matches_closed_form_and_is_reversible(bn, oracle_lib):
    """iid N(0,I), unit metric: one leapfrog is the linear map of src/kinetic_energy.jl:144-161."""
    Cn, D = 3, 17
    e = bn.Engine(Cn, D, lib=oracle_lib)
    e.model_iid_normal()
    rng = np.random.default_rng(5)
    q0 = rng.normal(size=(Cn, D)); p0 = rng.normal(size=(Cn, D))
    e.set_positions(q0)
    eps, n = 0.13, 7
    q, p = q0.copy(), p0.copy()
    for _ in range(n):
        pm = p + 0.5 * eps * (-q)
        q = q + eps * pm
        p = pm + 0.5 * eps * (-q)
    q1, p1, g1, l1 = e.leapfrog(p0, eps, n)
    np.testing.assert_allclose(q1, q, rtol=0, atol=1e-14)
    np.testing.assert_allclose(p1, p, rtol=0, atol=1e-14)
    np.testing.assert_allclose(g1, -q, rtol=0, atol=1e-14)
    np.testing.assert_allclose(l1, -0.5 * (q * q).sum(1), rtol=1e-14)
    # energy error is O(eps^2)
    H0 = -0.5 * (q0 * q0).sum(1) - 0.5 * (p0 * p0).sum(1)
    H1 = l1 - 0.5 * (p1 * p1).sum(1)
    e2 = bn.Engine(Cn, D, lib=oracle_lib); e2.model_iid_normal(); e2.set_positions(q0)
    _, ph, _, lh = e2.leapfrog(p0, eps / 2, 2 * n)
    Hh = lh - 0.5 * (ph * ph).sum(1)
    assert np.all(np.abs(Hh - H0) < 0.3 * np.abs(H1 - H0) + 1e-12)
    # reversibility: integrate back with -eps
    e.set_positions(q1)
    qb, pb, _, _ = e.leapfrog(p1, -eps, n)
    np.testing.assert_allclose(qb, q0, atol=1e-13)
    np.testing.assert_allclose(pb, p0, atol=1e-13)


def test_model_gradients_match_numpy(bn, oracle_lib):
    from conftest import make_logistic
    D = 9
    rng = np.random.default_rng(3)
    q = rng.normal(size=(2, D))
    # gaussian
    P, _ = make_gaussian(D)
    e = bn.Engine(2, D, lib=oracle_lib); e.model_gaussian(P); e.set_positions(q)
    _, g, l = e.get_state()
    np.testing.assert_allclose(g, -(q @ P.T), rtol=1e-12)
    np.testing.assert_allclose(l, -0.5 * np.einsum("cd,de,ce->c", q, P, q), rtol=1e-12)
    # funnel
    e = bn.Engine(2, D, lib=oracle_lib); e.model_funnel(); e.set_positions(q)
    _, g, l = e.get_state()
    v = q[:, 0]; S = (q[:, 1:] ** 2).sum(1)
    np.testing.assert_allclose(l, -v * v / 18 - 0.5 * (D - 1) * v - 0.5 * np.exp(-v) * S, rtol=1e-12)
    np.testing.assert_allclose(g[:, 0], -v / 9 - 0.5 * (D - 1) + 0.5 * np.exp(-v) * S, rtol=1e-12)
    np.testing.assert_allclose(g[:, 1:], -np.exp(-v)[:, None] * q[:, 1:], rtol=1e-12)
    # logistic, any row_blocks
    X, y, _ = make_logistic(257, D)
    for rb in (1, 5):
        e = bn.Engine(2, D, lib=oracle_lib); e.model_logistic(X, y, 0.7, row_blocks=rb); e.set_positions(q)
        _, g, l = e.get_state()
        eta = q @ X.T
        np.testing.assert_allclose(l, (y * eta - np.logaddexp(0, eta)).sum(1) - 0.35 * (q * q).sum(1), rtol=1e-12)
        np.testing.assert_allclose(g, (y - 1 / (1 + np.exp(-eta))) @ X - 0.7 * q, rtol=1e-10, atol=1e-11)


def test_dual_averaging_table(oracle_lib):
    """Fixed acceptance sequence through adapt_stepsize (src/stepsize.jl:208-229)."""
    class P(C.Structure):
        _fields_ = [("delta", C.c_double), ("gamma", C.c_double), ("kappa", C.c_double), ("t0", C.c_int32), ("_p", C.c_int32)]
    par = P(0.8, 0.05, 0.75, 10, 0)
    st = (C.c_double * 5)()
    eps0 = 0.37
    oracle_lib.bnuts_oracle_da_init.argtypes = [C.c_double, C.c_void_p]
    oracle_lib.bnuts_oracle_da_adapt.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    oracle_lib.bnuts_oracle_da_init(eps0, st)
    mu, m, Hbar, le, leb = math.log(10) + math.log(eps0), 0, 0.0, math.log(eps0), 0.0
    assert abs(st[0] - mu) < 1e-14 and st[1] == 0 and st[3] == pytest.approx(le, abs=1e-15) and st[4] == 0
    for a in [1.0, 0.3, 0.95, 0.0, 0.81, 0.6, 0.99, 0.2]:
        oracle_lib.bnuts_oracle_da_adapt(C.byref(par), st, a)
        m += 1
        Hbar += (0.8 - a - Hbar) / (m + 10)
        le = mu - math.sqrt(m) / 0.05 * Hbar
        leb += m ** (-0.75) * (le - leb)
        assert st[1] == m
        assert abs(st[2] - Hbar) < 1e-14 and abs(st[3] - le) < 1e-12 and abs(st[4] - leb) < 1e-12


def test_metric_update_matches_regularised_variance(oracle_lib):
    """src/hamiltonian.jl:153-162: M⁻¹ = var·N/(N+λ) + 1e-3·λ/(N+λ), W = 1/sqrt(M⁻¹)."""
    rng = np.random.default_rng(9)
    N, D = 57, 11
    x = rng.normal(size=(N, D)) * np.arange(1, D + 1) + 3.0
    lam = 5.0 / N
    minv = np.empty(D); w = np.empty(D)
    oracle_lib.bnuts_oracle_metric_update.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]
    oracle_lib.bnuts_oracle_metric_update(x.ctypes.data, D, N, D, lam, minv.ctypes.data, w.ctypes.data)
    want = x.var(axis=0, ddof=1) * N / (N + lam) + 1e-3 * lam / (N + lam)
    np.testing.assert_allclose(minv, want, rtol=1e-12)
    np.testing.assert_allclose(w, 1 / np.sqrt(want), rtol=1e-12)


def test_initial_stepsize_lands_in_band(bn, oracle_lib):
    """src/stepsize.jl:111-126: returned eps has A(eps) in [a_min, a_max] for the drawn momentum."""
    Cn, D = 16, 25
    e = bn.Engine(Cn, D, lib=oracle_lib, seed=99)
    e.model_funnel()
    e.set_positions(None)
    lib = _probe(oracle_lib)
    q, g, l = e.get_state()
    p = np.array([[lib.bnuts_oracle_normal(99, c, 0, d) for d in range(D)] for c in range(Cn)])
    e.find_initial_stepsize()
    eps = e.get_stepsize()
    _, p1, _, l1 = e.leapfrog(p, eps, 1)
    A = np.exp((l1 - 0.5 * (p1 * p1).sum(1)) - (l - 0.5 * (p * p).sum(1)))
    assert np.all((A >= 0.25 - 1e-9) & (A <= 0.75 + 1e-9))


def test_posterior_moments_within_mcse(bn, oracle_lib):
    Cn, D = 16, 20
    P, S = make_gaussian(D)
    e = bn.Engine(Cn, D, lib=oracle_lib, seed=4)
    e.model_gaussian(P)
    e.set_positions(None)
    e.find_initial_stepsize()
    for N, mk in [(75, 0), (25, 1), (50, 1), (100, 1), (200, 1), (50, 0)]:
        e.warmup_stage(N, mk, keep=False)
    ch, st = e.sample(600)
    x = ch.reshape(-1, D)
    sd = np.sqrt(np.diag(S))
    n_eff = x.shape[0] / 4.0   # conservative
    assert np.all(np.abs(x.mean(0)) < 5 * sd / np.sqrt(n_eff))
    assert np.all(np.abs(x.var(0) / np.diag(S) - 1) < 5 * np.sqrt(2 / n_eff))
    assert st["acceptance_rate"].mean() == pytest.approx(0.8, abs=0.1)

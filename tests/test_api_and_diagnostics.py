"""The host-side mirror of the reference API (api.py) and the diagnostics, driven on the CPU through the
host emulation of the product engine."""
import numpy as np
import pytest

from conftest import make_gaussian


def test_default_warmup_stages_match_reference(bn):
    """src/warmup.jl:366-371: optimum, search, 75 | 25,50,100,200,400 (diag) | 50 = 900 transitions."""
    st = bn.default_warmup_stages()
    assert isinstance(st[0], bn.FindLocalOptimum) and isinstance(st[1], bn.InitialStepsizeSearch)
    tn = [s for s in st if isinstance(s, bn.TuningNUTS)]
    assert [len(s) for s in tn] == [75, 25, 50, 100, 200, 400, 50]
    assert [s.M for s in tn] == [None] + ["Diagonal"] * 5 + [None]
    d = bn.DualAveraging()
    assert (d.δ, d.γ, d.κ, d.t0) == (0.8, 0.05, 0.75, 10)                       # src/stepsize.jl:191
    s = bn.InitialStepsizeSearch()
    assert (s.a_min, s.a_max, s.ϵ0, s.C, s.maxiter_crossing, s.maxiter_bisect) == (0.25, 0.75, 1.0, 2.0, 400, 400)
    n = bn.NUTS()
    assert (n.max_depth, n.min_Δ) == (10, -1000.0)                                # src/NUTS.jl:214


def test_threaded_mcmc_on_gaussian(bn, hostemu_lib):
    D = 6
    P, S = make_gaussian(D)
    stages = bn.default_warmup_stages(init_steps=40, middle_steps=25, doubling_stages=2, terminating_steps=30)
    chains, stats = bn.threaded_mcmc(bn.Gaussian(P), 300, nchains=8, warmup_stages=stages, lib=hostemu_lib, seed=5)
    assert chains.shape == (8, 300, D) and stats.shape == (8, 300)
    x = chains.reshape(-1, D)
    assert np.all(np.abs(x.mean(0)) < 6 * np.sqrt(np.diag(S) / 400))
    assert np.all(np.abs(x.var(0) / np.diag(S) - 1) < 0.35)
    s = bn.diagnostics.summarize_tree_statistics(stats)
    assert s.N == 2400 and 0.6 < s.a_mean < 0.95 and sum(s.termination_counts.values()) == 2400
    assert s.depth_counts.sum() == 2400
    eb = bn.diagnostics.EBFMI(stats)
    assert eb.shape == (8,) and np.all(eb > 0.3)
    c1, s1 = bn.mcmc_with_warmup(bn.IIDNormal(3), 50, warmup_stages=stages, lib=hostemu_lib)
    assert c1.shape == (50, 3) and s1.shape == (50,)


def test_ess_known_cases(bn):
    rng = np.random.default_rng(0)
    iid = rng.normal(size=(8, 2000))
    assert 0.8 * 16000 < bn.diagnostics.ess(iid) < 1.25 * 16000
    rho = 0.9                                        # AR(1): ESS = n (1 - rho) / (1 + rho)
    x = np.zeros((8, 4000))
    e = rng.normal(size=x.shape) * np.sqrt(1 - rho ** 2)
    for t in range(1, x.shape[1]):
        x[:, t] = rho * x[:, t - 1] + e[:, t]
    want = 8 * 4000 * (1 - rho) / (1 + rho)
    assert 0.7 * want < bn.diagnostics.ess(x) < 1.4 * want
    assert bn.diagnostics.min_ess(np.stack([iid, x[:, :2000]], axis=2)) < 0.2 * 16000


@pytest.mark.parametrize("dtype", [0, 1])
def test_fixed_stepsize_warmup_stages(bn, oracle_lib, hostemu_lib, dtype):
    """≙ fixed_stepsize_warmup_stages (src/warmup.jl:383-389) with FixedStepsize (src/stepsize.jl:251-255): the doubling
    metric windows run at the ϵ given in `initialization`; the step size never changes, the metric does; the host
    build of the product agrees with the oracle bit for bit."""
    stages = bn.fixed_stepsize_warmup_stages(local_optimization=None, middle_steps=10, doubling_stages=3)
    assert [s.N for s in stages[1:]] == [10, 20, 40] and all(isinstance(s.stepsize_adaptation, bn.FixedStepsize) for s in stages[1:])
    eps0 = np.array([0.11, 0.23, 0.35])
    outs = []
    for lib in (oracle_lib, hostemu_lib):
        r = bn.mcmc_keep_warmup(bn.Funnel(7), 12, warmup_stages=stages, nchains=3, lib=lib, seed=4, dtype=dtype,
                                initialization={"ϵ": eps0})
        for w in r["warmup"]:
            assert np.array_equal(w["warmup_state"]["ϵ"], r["initial_warmup_state"]["ϵ"])       # ϵ is kept ...
            assert np.all(w["results"][2] == r["initial_warmup_state"]["ϵ"][:, None])             # ... in every transition
        assert not np.array_equal(r["final_warmup_state"]["κ"], r["initial_warmup_state"]["κ"])   # the metric is tuned
        outs.append(r)
    a, b = outs
    for wa, wb in zip(a["warmup"], b["warmup"]):
        for x, y in zip(wa["results"], wb["results"]):
            assert np.asarray(x).tobytes() == np.asarray(y).tobytes()
    assert a["inference"][0].tobytes() == b["inference"][0].tobytes() and a["inference"][1].tobytes() == b["inference"][1].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [0, 1])
def test_cuda_fixed_stepsize_warmup_bitwise(bn, oracle_lib, cuda_lib, dtype):
    """The FixedStepsize stage on the CUDA engine (deterministic gradient path) against the oracle, bit for bit."""
    stages = bn.fixed_stepsize_warmup_stages(local_optimization=None, middle_steps=10, doubling_stages=3)
    runs = [bn.mcmc_keep_warmup(bn.Funnel(7), 12, warmup_stages=stages, nchains=3, lib=lib, seed=4, dtype=dtype,
                                initialization={"ϵ": np.array([0.11, 0.23, 0.35])}, **kw)
            for lib, kw in ((oracle_lib, {}), (cuda_lib, {"gradient_path": 1}))]
    a, b = runs
    for wa, wb in zip(a["warmup"], b["warmup"]):
        for x, y in zip(wa["results"], wb["results"]):
            assert np.asarray(x).tobytes() == np.asarray(y).tobytes()
    assert a["inference"][0].tobytes() == b["inference"][0].tobytes() and a["inference"][1].tobytes() == b["inference"][1].tobytes()

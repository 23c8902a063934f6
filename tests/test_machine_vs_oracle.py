"""The product's per-chain state machine, vector loops and host engine (compiled
for the host by tests/hostemu) against the recursive oracle: everything bit-for-bit."""
import numpy as np
import pytest

from conftest import run_protocol, assert_bitwise, set_model

F64, F32 = 0, 1


@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("kind", ["iid", "funnel", "gauss", "logit"])
def test_full_protocol_bitwise(bn, oracle_lib, hostemu_lib, kind, dtype):
    a = run_protocol(bn, oracle_lib, kind, dtype)
    b = run_protocol(bn, hostemu_lib, kind, dtype)
    assert_bitwise(a, b)


@pytest.mark.parametrize("dtype", [F64, F32])
def test_ragged_trees_divergences_and_max_depth(bn, oracle_lib, hostemu_lib, dtype):
    """Funnel with a coarse fixed step: deep ragged trees, divergences and max-depth hits."""
    outs = []
    for lib in (oracle_lib, hostemu_lib):
        e = bn.Engine(24, 10, dtype=dtype, max_depth=7, lib=lib, seed=5)
        e.model_funnel()
        e.set_positions(None)
        e.set_stepsize(np.linspace(0.05, 1.2, 24))
        ch, st, sel = e.sample(80, want_index=True)
        outs.append([ch, st, sel, e.get_state()[0]])
    assert_bitwise(outs[0], outs[1])
    st = outs[0][1]
    assert (st["term_left"] == st["term_right"]).any(), "no divergence exercised"
    assert ((st["term_left"] == 1) & (st["term_right"] == 0)).any(), "max depth never reached"
    assert ((st["term_left"] < st["term_right"])).any(), "no turning termination"
    assert len(np.unique(st["depth"])) >= 4


def test_injected_momenta_and_directions(bn, oracle_lib, hostemu_lib):
    rng = np.random.default_rng(2)
    Cn, D, T = 5, 33, 7
    p = rng.normal(size=(T, Cn, D))
    dirs = rng.integers(0, 2 ** 32, size=(T, Cn), dtype=np.uint64).astype(np.uint32)
    outs = []
    for lib in (oracle_lib, hostemu_lib):
        e = bn.Engine(Cn, D, max_depth=6, lib=lib)
        set_model(e, "gauss", D)
        e.set_positions(rng.normal(size=(Cn, D)) * 0 + 0.3)
        e.set_stepsize(0.21)
        e.inject(T, dirs, p)
        outs.append(list(e.sample(T + 3, want_index=True)))   # last 3 transitions fall back to Philox
    assert_bitwise(outs[0], outs[1])


def test_nonfinite_start_is_reported_per_chain(bn, oracle_lib, hostemu_lib):
    for lib in (oracle_lib, hostemu_lib):
        e = bn.Engine(3, 4, lib=lib)
        e.model_funnel()
        q = np.zeros((3, 4)); q[1, 0] = -800.0; q[1, 1] = 1e200   # exp(800)*1e400 -> non-finite
        rc = e.set_positions(q, allow_nonfinite=True)
        assert rc == -4
        assert list(e.chain_status()) == [0, -4, 0]
        with pytest.raises(bn.BnutsError):
            e.set_positions(q)


def test_chain_offset_makes_sharding_invisible(bn, hostemu_lib):
    """Chains are keyed by global id: two engines of 3 chains == one engine of 6."""
    def run(C, off):
        e = bn.Engine(C, 8, max_depth=5, lib=hostemu_lib, seed=3, chain_offset=off)
        e.model_funnel(); e.set_positions(None); e.set_stepsize(0.3)
        return e.sample(20)
    full = run(6, 0)
    a, b = run(3, 0), run(3, 3)
    assert np.concatenate([a[0], b[0]]).tobytes() == full[0].tobytes()
    assert np.concatenate([a[1], b[1]]).tobytes() == full[1].tobytes()


@pytest.mark.parametrize("C,D,max_depth,eps,kind", [
    (1, 1, 3, 0.4, "iid"),          # one chain, one coordinate
    (2, 1, 1, 0.9, "funnel"),       # max_depth 1: every tree is a single doubling; funnel with D = 1 is N(0, 9)
    (3, 32, 2, 0.2, "gauss"),       # D exactly one lane row
    (3, 33, 4, 0.2, "gauss"),       # one element into the next group of 32
    (2, 128, 3, 0.05, "logit"),     # D = 128: four full lane rows
    (2, 129, 3, 0.05, "logit"),
    (4, 5, 20, 1e-3, "iid"),        # large max_depth, tiny step: deep trees (capped by the draw count)
    (3, 4, 32, 0.02, "iid"),        # max_depth at the limit (≙ MAX_DIRECTIONS_DEPTH = 32, src/tree.jl:132): 36 phase-point slots
    (4, 7, 6, 50.0, "funnel"),      # absurd step: every first leaf diverges
    (5, 6, 5, 1e-7, "iid"),         # near-zero step: every tree hits max depth without turning
])
def test_edge_shapes_and_regimes_bitwise(bn, oracle_lib, hostemu_lib, C, D, max_depth, eps, kind):
    """Extreme shapes and integrator regimes, the host build of the product state machine against the oracle:
    draws, statistics, selected indices and final state bit for bit; the regime each case is meant to reach is
    asserted so the test cannot pass vacuously."""
    outs = []
    n = 6 if max_depth >= 20 else 25
    for lib in (oracle_lib, hostemu_lib):
        e = bn.Engine(C, D, max_depth=max_depth, lib=lib, seed=13)
        set_model(e, kind, D, N=120)
        e.set_positions(None)
        e.set_stepsize(eps)
        ch, st, sel = e.sample(n, want_index=True)
        one = e.sample(1)                                      # a single-draw call
        outs.append([ch, st, sel, one[0], one[1], e.get_state()[0], e.get_state()[1]])
    assert_bitwise(outs[0], outs[1])
    st = outs[0][1]
    assert (st["depth"] <= max_depth).all() and (st["steps"] >= 1).all()
    if eps >= 50.0:
        assert (st["term_left"] == st["term_right"]).mean() > 0.9          # divergence at a leaf: InvalidTree(i, i)
        assert (outs[0][0][:, 0] == outs[0][0][:, -1]).all()               # nothing is ever accepted: chains stay put
    if eps <= 1e-7:
        assert ((st["term_left"] == 1) & (st["term_right"] == 0)).all()    # REACHED_MAX_DEPTH every time
        assert (st["steps"] == 2 ** max_depth - 1).all()
    if max_depth == 1:
        assert (st["steps"] == 1).all()


def test_config_c1_full_run_bitwise_and_moments(bn, oracle_lib, hostemu_lib):
    """BASELINE config 1 exactly as stated (100-dim iid standard normal, diagonal metric, max depth 10, 1000 warmup
    (75 | 25..400 | 150) + 1000 draws, one chain): the host build of the product against the oracle bit for bit over
    the whole run, and the draws against N(0, I) within Monte-Carlo error."""
    stages = bn.default_warmup_stages(local_optimization=None, terminating_steps=150)
    runs = [bn.mcmc_keep_warmup(bn.IIDNormal(100), 1000, warmup_stages=stages, nchains=1, lib=lib, seed=1)
            for lib in (oracle_lib, hostemu_lib)]
    a, b = runs
    assert sum(w["stage"].N for w in a["warmup"] if hasattr(w["stage"], "N")) == 1000
    for wa, wb in zip(a["warmup"], b["warmup"]):
        for k in ("q", "κ", "W", "ϵ"):
            assert wa["warmup_state"][k].tobytes() == wb["warmup_state"][k].tobytes()
    assert a["inference"][0].tobytes() == b["inference"][0].tobytes() and a["inference"][1].tobytes() == b["inference"][1].tobytes()
    ch, st = a["inference"]
    ess = np.array([bn.diagnostics.ess(ch[:, :, d], rank_normalise=False) for d in range(100)])   # one chain per coordinate
    ess2 = np.array([bn.diagnostics.ess(ch[:, :, d] ** 2, rank_normalise=False) for d in range(100)])   # of the second moment
    assert np.all(np.abs(ch[0].mean(0)) < 5 / np.sqrt(ess)) and np.all(np.abs((ch[0] ** 2).mean(0) - 1) < 5 * np.sqrt(2 / ess2))
    assert (st["depth"] <= 10).all() and (st["term_left"] != st["term_right"]).all()      # no divergences on N(0, I)

"""Randomised differential test: the host build of the product state machine against the recursive oracle on
small random configurations (shape, target, depth cap, step size, divergence threshold, seed, dtype) — draws,
statistics, selected indices, adapted step size and metric bit for bit.  Deterministic (derandomised hypothesis)."""
import numpy as np
from hypothesis import given, settings, strategies as st, HealthCheck

from conftest import set_model


@settings(max_examples=1000, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(C=st.integers(1, 5), D=st.integers(1, 70), depth=st.integers(1, 8), kind=st.sampled_from(["iid", "funnel", "gauss", "logit"]),
       eps=st.floats(1e-3, 3.0), min_delta=st.sampled_from([-1000.0, -5.0, -0.5]), seed=st.integers(0, 2 ** 40),
       dtype=st.sampled_from([0, 1]), adapt=st.booleans())
def test_host_build_equals_oracle(bn, oracle_lib, hostemu_lib, C, D, depth, kind, eps, min_delta, seed, dtype, adapt):
    outs = []
    for lib in (oracle_lib, hostemu_lib):
        e = bn.Engine(C, D, dtype=dtype, max_depth=depth, min_delta=min_delta, lib=lib, seed=seed)
        set_model(e, kind, D, seed=seed % 1000 + 1, N=60)
        e.set_positions(None)
        e.set_stepsize(eps)
        o = []
        if adapt:
            o += list(e.warmup_stage(12, 1, allow_fail=True) or ())
        ch, stt, sel = e.sample(8, want_index=True)
        o += [ch, stt, sel, e.get_stepsize(), e.get_metric_diag(), e.get_metric_diag_w(), np.array(e.chain_status())]
        outs.append(o)
        e.close()
    assert len(outs[0]) == len(outs[1])
    for x, y in zip(outs[0], outs[1]):
        assert np.asarray(x).tobytes() == np.asarray(y).tobytes()

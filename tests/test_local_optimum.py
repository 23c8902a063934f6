"""≙ warmup!(FindLocalOptimum), src/warmup.jl:152-186: the batched ascent of the device state machine against the
oracle's independent restatement (bit for bit), against the known optimum, and the restart rule."""
import numpy as np
import pytest

from conftest import set_model, make_logistic, make_gaussian, assert_bitwise

F64, F32 = 0, 1


def _run(bn, lib, kind, dtype, C=6, D=11, penalty=1e-4, iters=50, **kw):
    e = bn.Engine(C, D, dtype=dtype, lib=lib, seed=9, **kw)
    set_model(e, kind, D, N=200)
    e.set_positions(None)
    e.find_local_optimum(penalty, iters)
    out = list(e.get_state())
    e.close()
    return out


@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("kind", ["iid", "funnel", "gauss", "logit"])
def test_machine_matches_oracle_bitwise_cpu(bn, oracle_lib, hostemu_lib, kind, dtype):
    assert_bitwise(_run(bn, oracle_lib, kind, dtype), _run(bn, hostemu_lib, kind, dtype))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("kind", ["iid", "funnel", "gauss", "logit"])
def test_machine_matches_oracle_bitwise_cuda(bn, oracle_lib, cuda_lib, kind, dtype):
    assert_bitwise(_run(bn, oracle_lib, kind, dtype), _run(bn, cuda_lib, kind, dtype, gradient_path=1))


def test_reaches_the_known_optimum(bn, oracle_lib, hostemu_lib):
    D, C = 9, 5
    # Gaussian: argmax of -½ qᵀPq - ½λ‖q‖² is 0
    for lib in (oracle_lib, hostemu_lib):
        e = bn.Engine(C, D, lib=lib, seed=3); e.model_gaussian(make_gaussian(D)[0]); e.set_positions(None)
        l0 = e.get_state()[2]
        e.find_local_optimum(1e-4, 200)
        q, g, l = e.get_state()
        assert np.all(l >= l0) and np.max(np.abs(q)) < 1e-6 and np.max(np.abs(g)) < 1e-5
        e.close()
    # logistic regression: stationarity of the penalised objective, and agreement with Newton on the host
    N = 500
    X, y, beta = make_logistic(N, D, seed=4)
    e = bn.Engine(C, D, lib=hostemu_lib, seed=3); e.model_logistic(X, y, 1.0); e.set_positions(None)
    e.find_local_optimum(1e-4, 300)
    q, g, l = e.get_state()
    b = np.zeros(D)
    for _ in range(30):
        s = 1 / (1 + np.exp(-X @ b))
        b += np.linalg.solve((X * (s * (1 - s))[:, None]).T @ X + (1 + 1e-4) * np.eye(D), X.T @ (y - s) - (1 + 1e-4) * b)
    np.testing.assert_allclose(q, np.tile(b, (C, 1)), atol=1e-5)
    assert np.max(np.abs(g - 1e-4 * q)) < 1e-4
    # a few iterations only: still an improvement ("we don't need to find the mode", src/warmup.jl:146-147)
    e.set_positions(None); l0 = e.get_state()[2]; e.find_local_optimum(1e-4, 3)
    assert np.all(e.get_state()[2] > l0)
    e.close()


def test_nonfinite_start_is_rerandomised(bn, oracle_lib, hostemu_lib):
    """≙ src/warmup.jl:162-172: a non-finite start draws a new position (and doubles the penalty)."""
    D, C = 4, 3
    outs = []
    for lib in (oracle_lib, hostemu_lib):
        e = bn.Engine(C, D, lib=lib, seed=5); e.model_funnel()
        q0 = np.zeros((C, D)); q0[1, 0] = -800.0; q0[1, 1] = 1e200      # exp(800) * 1e400 overflows: ℓ = -Inf
        e.set_positions(q0, allow_nonfinite=True)
        assert not np.isfinite(e.get_state()[2][1])
        e.find_local_optimum(1e-4, 20)
        q, g, l = e.get_state()
        assert np.all(np.isfinite(l)) and np.all(np.abs(q[1]) < 50) and (e.chain_status() == 0).all()
        outs.append([q, g, l]); e.close()
    assert_bitwise(outs[0], outs[1])

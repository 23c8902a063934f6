"""Error behaviour of the C ABI (include/bnuts.h): every call returns 0 or a negative bnuts_status, nothing throws
across the boundary, numerical failures are per chain (≙ the exceptions of src/warmup.jl:151,172,291-296,
src/stepsize.jl:71,101,128).  Runs on the oracle and on the host build of the product engine."""
import ctypes as C

import numpy as np
import pytest

from conftest import make_gaussian, make_logistic


@pytest.fixture(params=["oracle", "hostemu"])
def lib(request, oracle_lib, hostemu_lib):
    return oracle_lib if request.param == "oracle" else hostemu_lib


def test_create_rejects_bad_config(bn, lib):
    for kw in (dict(C=0, D=3), dict(C=2, D=0), dict(C=2, D=3, max_depth=0), dict(C=2, D=3, max_depth=33),
               dict(C=2, D=3, min_delta=1.0)):
        with pytest.raises(bn.BnutsError) as ei:
            bn.Engine(kw.pop("C"), kw.pop("D"), lib=lib, **kw)
        assert ei.value.code == -1
    assert lib.bnuts_destroy(None) == 0 and lib.bnuts_set_positions(None, None) == -1


def test_calls_without_a_model(bn, lib):
    e = bn.Engine(2, 3, lib=lib)
    for call in (lambda: e.set_positions(None), lambda: e.sample(2), lambda: e.find_initial_stepsize(),
                 lambda: e.find_local_optimum(), lambda: e.leapfrog(np.zeros((2, 3)), 0.1, 1)):
        with pytest.raises(bn.BnutsError) as ei:
            call()
        assert ei.value.code == -2
    e.close()


def test_invalid_arguments(bn, lib):
    e = bn.Engine(3, 4, lib=lib); e.model_iid_normal(); e.set_positions(None)
    with pytest.raises(bn.BnutsError) as ei:
        e.sample(0)
    assert ei.value.code == -1
    with pytest.raises(bn.BnutsError) as ei:
        e.find_local_optimum(-1.0, 10)
    assert ei.value.code == -1
    bad = np.eye(4); bad[0, 0] = -1.0
    with pytest.raises(bn.BnutsError) as ei:
        e.set_metric_dense(bad)
    assert ei.value.code == -1
    # the invariants the reference records as commented-out @argcheck's (src/stepsize.jl:31-35,183-186)
    for kw in (dict(delta=0.0), dict(delta=1.0), dict(gamma=0.0), dict(kappa=0.5), dict(kappa=1.5), dict(t0=-1)):
        with pytest.raises(bn.BnutsError) as ei:
            e.warmup_stage(5, 0, **kw)
        assert ei.value.code == -1, kw
    for args in ((0.0, 0.75), (0.8, 0.75), (0.25, 1.0)):
        with pytest.raises(bn.BnutsError) as ei:
            e.find_initial_stepsize(*args)
        assert ei.value.code == -1, args
    with pytest.raises(bn.BnutsError) as ei:
        e.find_initial_stepsize(0.25, 0.75, 1.0, 1.0)            # C must exceed 1
    assert ei.value.code == -1
    e.set_stepsize(0.3); e.warmup_stage(5, 0, fixed_stepsize=True)   # FixedStepsize needs no adaptation parameters
    assert e.chain_status().tolist() == [0, 0, 0]
    e.close()


def test_unsupported_requests_fail_loudly(bn, hostemu_lib):
    """No silent fallbacks: the tensor path (D <= 256), its reference point and the NCCL / peer-memory exchanges are CUDA-only."""
    X, y, _ = make_logistic(50, 4)
    e = bn.Engine(2, 4, dtype=bn.F32, lib=hostemu_lib, gradient_path=bn.GRAD_TENSOR)
    with pytest.raises(bn.BnutsError) as ei:
        e.model_logistic(X, y, 1.0)
    assert ei.value.code == -7
    e.close()
    e = bn.Engine(2, 4, lib=hostemu_lib); e.model_logistic(X, y, 1.0)
    for call in (lambda: e.logistic_set_reference(np.zeros(4)), lambda: e.set_nccl(bytes(128), 1, 0), lambda: e.p2p_export()):
        with pytest.raises(bn.BnutsError) as ei:
            call()
        assert ei.value.code == -7
    e.model_funnel()
    with pytest.raises(bn.BnutsError) as ei:
        e.set_metric_dense(np.eye(4))
    assert ei.value.code == -7
    e.close()


def test_nonfinite_start_is_per_chain(bn, lib):
    e = bn.Engine(3, 2, lib=lib); e.model_funnel()
    q = np.zeros((3, 2)); q[1] = [-800.0, 1e200]
    rc = e.set_positions(q, allow_nonfinite=True)
    assert rc == -4 and e.chain_status().tolist() == [0, -4, 0]
    assert b"non-finite" in lib.bnuts_last_error(e.h)
    e.set_stepsize(0.1)
    ch, st = e.sample(3)                               # the healthy chains still run
    assert (st["steps"][[0, 2]] > 0).all()
    e.close()


def test_stepsize_collapse_is_reported(bn, lib):
    """≙ the ϵ < 1e-10 assertion of src/warmup.jl:291-296: reported as a status, the batch is not aborted."""
    P, _ = make_gaussian(3)
    e = bn.Engine(2, 3, lib=lib); e.model_gaussian(P * 1e30); e.set_positions(np.ones((2, 3)))
    e.set_stepsize(1e-9)
    rc = e.warmup_stage(60, 0, allow_fail=True)
    st = e.chain_status()
    assert (rc is None or isinstance(rc, tuple) or rc in (0, -6)) and set(st.tolist()) <= {0, -6}
    e.close()

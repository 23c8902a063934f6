"""Synthetic design matrix of the row-sharded configuration (include/bnuts.h: bnuts_model_logistic_synthetic,
bnuts_synth_logistic_rows).  The reference ships no data: the definition is this
project's own (bnuts_math.h, synth_*), pinned here by a frozen fixture, structural properties and statistics; the
oracle and the host build of the product code must agree bit for bit, and so must any sharding of the rows."""
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "synth_rows.npz")
SEED, D = 5, 10
BLOCKS = ((0, 6), (4294967294, 4), (12500000 * 7 + 3, 3))


def _bf16(bits):
    return (bits.astype(np.uint32) << 16).view(np.float32).astype(np.float64)


@pytest.mark.parametrize("which", ["oracle", "hostemu"])
def test_frozen_rows(bn, oracle_lib, hostemu_lib, which):
    lib = oracle_lib if which == "oracle" else hostemu_lib
    z = np.load(GOLD)
    for k, (r0, n) in enumerate(BLOCKS):
        X, y, beta = bn.synth_logistic_rows(SEED, r0, n, D, lib=lib)
        assert X.tobytes() == z[f"X{k}"].tobytes() and y.tobytes() == z[f"y{k}"].tobytes()
    assert beta.tobytes() == z["beta"].tobytes()


def test_sharding_invariance_and_structure(bn, oracle_lib, hostemu_lib):
    N, Dd = 5000, 37                                  # D not a multiple of 4: the last quad is partial
    X, y, beta = bn.synth_logistic_rows(7, 1000, N, Dd, lib=oracle_lib)
    Xh, yh, bh = bn.synth_logistic_rows(7, 1000, N, Dd, lib=hostemu_lib)
    assert X.tobytes() == Xh.tobytes() and y.tobytes() == yh.tobytes() and beta.tobytes() == bh.tobytes()
    parts = [bn.synth_logistic_rows(7, 1000 + a, b - a, Dd, lib=oracle_lib) for a, b in ((0, 1), (1, 1777), (1777, N))]
    assert np.concatenate([p[0] for p in parts]).tobytes() == X.tobytes()
    assert np.concatenate([p[1] for p in parts]).tobytes() == y.tobytes()
    x = _bf16(X)
    assert np.all(x[:, 0] == 1.0) and set(np.unique(y)) <= {0.0, 1.0}
    # a different seed or a different row gives different data; the same (seed, row) the same
    X2, _, _ = bn.synth_logistic_rows(8, 1000, 4, Dd, lib=oracle_lib)
    assert X2.tobytes() != X[:4].tobytes()
    # the label rule, recomputed with numpy Float64: P(y = 1) = sigma(x . beta*); the labels are consistent with it
    p = 1 / (1 + np.exp(-(x @ beta)))
    assert abs(y.mean() - p.mean()) < 4 * np.sqrt(0.25 / N)
    assert np.corrcoef(y, p)[0, 1] > 0.2
    # columns 1.. are N(0,1) rounded to bf16: moments within sampling error, no duplicated columns / rows
    zc = x[:, 1:]
    assert abs(zc.mean()) < 4 / np.sqrt(zc.size) and abs(zc.var() - 1) < 0.02
    assert np.max(np.abs(np.corrcoef(zc.T) - np.eye(Dd - 1))) < 0.08
    assert abs(np.mean(beta ** 2) * Dd - 1) < 0.7      # beta* ~ N(0, 1/D)


@pytest.mark.parametrize("which", ["oracle", "hostemu"])
def test_synthetic_model_equals_explicit_model(bn, oracle_lib, hostemu_lib, which):
    """bnuts_model_logistic_synthetic == bnuts_model_logistic on the rows returned by bnuts_synth_logistic_rows."""
    lib = oracle_lib if which == "oracle" else hostemu_lib
    N, Dd, C = 700, 12, 6
    X, y, beta = bn.synth_logistic_rows(11, 123456789012, N, Dd, lib=lib)
    rng = np.random.default_rng(0)
    q = beta[None, :] + rng.normal(size=(C, Dd)) * 0.3
    a = bn.Engine(C, Dd, dtype=bn.F64, lib=lib); a.model_logistic_synthetic(11, 123456789012, N, 1.0, row_blocks=2)
    b = bn.Engine(C, Dd, dtype=bn.F64, lib=lib); b.model_logistic(X, y, 1.0, row_blocks=2)
    a.set_positions(q); b.set_positions(q)
    for u, v in zip(a.get_state(), b.get_state()):
        assert u.tobytes() == v.tobytes()
    with pytest.raises(bn.BnutsError):
        a.model_logistic_synthetic(11, -1, N)

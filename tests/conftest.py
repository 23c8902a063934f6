"""Shared fixtures.  `-m "not gpu"` runs on CPU only; `-m gpu` needs a B200."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ORACLE_SO = os.path.join(ROOT, "oracle", "libbnuts_oracle.so")
HOSTEMU_SO = os.path.join(ROOT, "tests", "hostemu", "libbnuts_hostemu.so")
CUDA_SO = os.path.join(ROOT, "inplacedhmc.jl_b200", "csrc", "libbnuts.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _csrc(*names):
    return [os.path.join(ROOT, "inplacedhmc.jl_b200", "csrc", n) for n in names]


def build_oracle():
    src = [os.path.join(ROOT, "oracle", "bnuts_oracle.cpp"), os.path.join(ROOT, "include", "bnuts.h")] + \
        _csrc("bnuts_math.h", "bnuts_models.h")
    if not _newer(ORACLE_SO, src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return ORACLE_SO


def build_hostemu():
    d = os.path.join(ROOT, "tests", "hostemu")
    src = [os.path.join(d, "hostemu.cpp")] + _csrc("bnuts_math.h", "bnuts_models.h", "nuts_machine.h", "backend.h",
                                                    "engine_core.h", "capi_impl.h")
    if not _newer(HOSTEMU_SO, src):
        subprocess.check_call(["make", "-s", "-C", d])
    return HOSTEMU_SO


@pytest.fixture(scope="session")
def bn():
    import inplacedhmc_jl_b200 as mod
    return mod


@pytest.fixture(scope="session")
def oracle_lib(bn):
    return bn.load_library(build_oracle())


@pytest.fixture(scope="session")
def hostemu_lib(bn):
    return bn.load_library(build_hostemu())


@pytest.fixture(scope="session")
def cuda_lib(bn):
    import torch  # noqa: F401  (device plumbing only)
    assert torch.cuda.is_available(), "gpu test without a CUDA device"
    return bn.load_library(CUDA_SO)  # raises ImportError loudly if the extension was not built


# ---------------------------------------------------------------- synthetic problems
def make_gaussian(D, seed=2):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(D, 2 * D))
    S = A @ A.T / (2 * D)
    return np.linalg.inv(S), S


def to_bf16_grid(x):
    """round float64 -> nearest bf16 value (ties to even), returned as float64"""
    f = np.asarray(x, dtype=np.float32)
    u = f.view(np.uint32)
    r = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return r.astype(np.uint32).view(np.float32).astype(np.float64)


def make_logistic(N, D, seed=3):
    rng = np.random.default_rng(seed)
    X = to_bf16_grid(rng.normal(size=(N, D)))
    X[:, 0] = 1.0
    beta = rng.normal(size=D) / np.sqrt(D)
    y = (rng.uniform(size=N) < 1.0 / (1.0 + np.exp(-X @ beta))).astype(np.float64)
    return X, y, beta


def set_model(e, kind, D, seed=1, N=300, row_blocks=4):
    if kind == "iid":
        e.model_iid_normal()
    elif kind == "funnel":
        e.model_funnel()
    elif kind == "gauss":
        e.model_gaussian(make_gaussian(D, seed)[0])
    elif kind == "logit":
        X, y, _ = make_logistic(N, D, seed)
        e.model_logistic(X, y, 1.0, row_blocks=row_blocks)
    else:
        raise ValueError(kind)


def run_protocol(bn, lib, kind, dtype, C=6, D=37, max_depth=6, seed=7, stages=((30, 0), (25, 1), (40, 1), (20, 0)),
                 n_draws=60, **kw):
    """A fixed end-to-end script (search, windowed warmup, draws, bare leapfrogs);
    returns a flat list of arrays for bit-for-bit comparison between libraries."""
    e = bn.Engine(C, D, dtype=dtype, max_depth=max_depth, lib=lib, seed=seed, **kw)
    set_model(e, kind, D)
    e.set_positions(None)
    out = list(e.get_state())
    e.find_initial_stepsize()
    out.append(e.get_stepsize())
    for N, mk in stages:
        ch, st, ep = e.warmup_stage(N, mk)
        out += [ch, st, ep]
    out += [e.get_metric_diag(), e.get_stepsize()]
    ch, st, sel = e.sample(n_draws, want_index=True)
    out += [ch, st, sel]
    rng = np.random.default_rng(11)
    p = rng.normal(size=(C, D))
    out += list(e.leapfrog(p, 0.1, 3)) + list(e.leapfrog(p, -0.05, 2))
    c = e.counters()
    out.append(np.array([c["leapfrogs"], c["transitions"], c["divergences"]]))
    e.close()
    return out


def assert_bitwise(a, b, names=None):
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        x = np.asarray(x); y = np.asarray(y)
        assert x.shape == y.shape and x.dtype == y.dtype, (i, x.shape, y.shape)
        assert x.tobytes() == y.tobytes(), f"output {i} differs"

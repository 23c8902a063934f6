"""Host-side mirror of InplaceDHMC.jl's user API for the NUTS hot path.

Same names, argument meaning and defaults as the reference (export list
src/InplaceDHMC.jl:3-11); the bodies drive the C ABI of include/bnuts.h instead of
Julia code.  The Julia twin of this file is julia/BNuts.jl (ccall bindings).

    chains, stats = threaded_mcmc(model, 1000; nchains = 4096)

`model` is one of the built-in targets (the reference takes a user
AbstractProbabilityModel; a device engine cannot call a host closure):
IIDNormal(D), Funnel(D), Gaussian(P), Logistic(X, y, prior_precision).
"""
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _capi as capi

DEFAULT_MAX_TREE_DEPTH = 10


# ---------------------------------------------------------------- models
@dataclass
class IIDNormal:
    dim: int

    def attach(self, e):
        e.model_iid_normal()


@dataclass
class Funnel:
    dim: int

    def attach(self, e):
        e.model_funnel()


@dataclass
class Gaussian:
    precision: np.ndarray

    @property
    def dim(self):
        return self.precision.shape[0]

    def attach(self, e):
        e.model_gaussian(self.precision)


@dataclass
class Logistic:
    X: np.ndarray
    y: np.ndarray
    prior_precision: float = 1.0
    row_blocks: int = 1

    @property
    def dim(self):
        return self.X.shape[1]

    def attach(self, e):
        e.model_logistic(self.X, self.y, self.prior_precision, self.row_blocks)


# ---------------------------------------------------------------- options (same defaults as the reference)
@dataclass
class NUTS:
    """≙ NUTS(; max_depth, min_Δ), src/NUTS.jl:204-220"""
    max_depth: int = DEFAULT_MAX_TREE_DEPTH
    min_Δ: float = -1000.0


@dataclass
class DualAveraging:
    """≙ DualAveraging(; δ, γ, κ, t₀), src/stepsize.jl:173-193"""
    δ: float = 0.8
    γ: float = 0.05
    κ: float = 0.75
    t0: int = 10


@dataclass
class FixedStepsize:
    """≙ FixedStepsize, src/stepsize.jl:251-255: a warmup stage with this adaptation keeps ϵ."""


@dataclass
class InitialStepsizeSearch:
    """≙ InitialStepsizeSearch(; a_min, a_max, ϵ₀, C, maxiter_crossing, maxiter_bisect), src/stepsize.jl:16-38"""
    a_min: float = 0.25
    a_max: float = 0.75
    ϵ0: float = 1.0
    C: float = 2.0
    maxiter_crossing: int = 400
    maxiter_bisect: int = 400


@dataclass
class FindLocalOptimum:
    """≙ src/warmup.jl:137-150 -> bnuts_find_local_optimum.  The reference's inner solver is the un-vendored
    QuasiNewtonMethods.proptimize!; the engine climbs with a Barzilai-Borwein step (see include/bnuts.h)."""
    magnitude_penalty: float = 1e-4
    iterations: int = 50


@dataclass
class TuningNUTS:
    """≙ TuningNUTS{M}(N, stepsize_adaptation, λ), src/warmup.jl:217-234.
    M is "Diagonal" or None (the reference accepts Symmetric but still adapts a diagonal)."""
    N: int
    stepsize_adaptation: DualAveraging = field(default_factory=DualAveraging)
    M: Optional[str] = "Diagonal"
    λ: Optional[float] = None   # default 5/N

    def __len__(self):
        return self.N


@dataclass
class GaussianKineticEnergy:
    """≙ GaussianKineticEnergy (diagonal M⁻¹, W = 1/sqrt(M⁻¹)), src/hamiltonian.jl:33-74"""
    M_inv: np.ndarray

    @property
    def W(self):
        return 1.0 / np.sqrt(self.M_inv)


class NoProgressReport:
    """≙ src/reporting.jl:6"""


class LogProgressReport:
    """≙ src/reporting.jl:39-46 (host-side logging only; not part of the hot path)"""

    def __init__(self, step_interval=100):
        self.step_interval = step_interval


def default_warmup_stages(local_optimization=FindLocalOptimum(), stepsize_search=InitialStepsizeSearch(), M="Diagonal",
                          stepsize_adaptation=None, init_steps=75, middle_steps=25, doubling_stages=5,
                          terminating_steps=50):
    """≙ default_warmup_stages, src/warmup.jl:361-372: 75 | 25,50,100,200,400 | 50."""
    da = stepsize_adaptation or DualAveraging()
    mid = tuple(TuningNUTS(middle_steps << d, da, M) for d in range(doubling_stages))
    return (local_optimization, stepsize_search, TuningNUTS(init_steps, da, None)) + mid + \
        (TuningNUTS(terminating_steps, da, None),)


def fixed_stepsize_warmup_stages(local_optimization=FindLocalOptimum(), M="Diagonal", middle_steps=25, doubling_stages=5):
    """≙ fixed_stepsize_warmup_stages, src/warmup.jl:383-389: local optimisation, then the doubling metric windows with the
    step size fixed at the ϵ given in `initialization` (no step-size search, no dual averaging)."""
    return (local_optimization,) + tuple(TuningNUTS(middle_steps << d, FixedStepsize(), M) for d in range(doubling_stages))


def _run_warmup(e, stages):
    for st in stages:
        if st is None:
            continue
        if isinstance(st, FindLocalOptimum):
            e.find_local_optimum(st.magnitude_penalty, st.iterations)
        elif isinstance(st, InitialStepsizeSearch):
            e.find_initial_stepsize(st.a_min, st.a_max, st.ϵ0, st.C, st.maxiter_crossing, st.maxiter_bisect)
        elif isinstance(st, TuningNUTS):
            da = st.stepsize_adaptation
            fixed = isinstance(da, FixedStepsize)
            da = DualAveraging() if fixed else da
            e.warmup_stage(st.N, capi.METRIC_DIAG if st.M else capi.METRIC_NONE, da.δ, da.γ, da.κ, da.t0,
                           -1.0 if st.λ is None else st.λ, keep=False, fixed_stepsize=fixed)
        else:
            raise TypeError(f"unknown warmup stage {st!r}")


def threaded_mcmc(ℓ, N, δ=0.8, initialization=None, warmup_stages=None, algorithm=None, reporter=None, nchains=4096,
                  dtype=capi.F64, seed=20261018, device=0, chain_offset=0, gradient_path=capi.GRAD_AUTO, lib=None,
                  return_engine=False):
    """≙ threaded_mcmc(ℓ, N; δ, initialization, warmup_stages, algorithm, reporter, nchains), src/mcmc.jl:130-159.

    Returns (chains [nchains, N, D] Float64, tree_statistics [nchains, N] of TreeStatisticsNUTS records).
    `initialization` may carry q [nchains, D], κ (GaussianKineticEnergy with per-chain or shared diagonal) and ϵ.
    Deviation from the reference: a given ϵ REPLACES the InitialStepsizeSearch stage (the reference, whose argument check
    is commented out, would still run the search and overwrite the value, src/warmup.jl:188-200); julia/BNuts.jl does the same.
    """
    algorithm = algorithm or NUTS()
    initialization = initialization or {}
    if warmup_stages is None:
        warmup_stages = default_warmup_stages(stepsize_adaptation=DualAveraging(δ=δ))
    D = ℓ.dim
    e = capi.Engine(nchains, D, dtype=dtype, max_depth=algorithm.max_depth, min_delta=algorithm.min_Δ, seed=seed,
                    chain_offset=chain_offset, device=device, gradient_path=gradient_path, lib=lib)
    ℓ.attach(e)
    κ = initialization.get("κ")
    if κ is not None:
        e.set_metric_diag(np.broadcast_to(np.asarray(κ.M_inv, dtype=np.float64), (nchains, D)))
    e.set_positions(initialization.get("q"))
    if initialization.get("ϵ") is not None:
        e.set_stepsize(initialization["ϵ"])
        warmup_stages = tuple(s for s in warmup_stages if not isinstance(s, InitialStepsizeSearch))
    _run_warmup(e, warmup_stages)
    chains, stats = e.sample(N)
    if return_engine:
        return chains, stats, e
    e.close()
    return chains, stats


def mcmc_keep_warmup(ℓ, N, δ=0.8, initialization=None, warmup_stages=None, algorithm=None, nchains=4096, dtype=capi.F64,
                     seed=20261018, device=0, chain_offset=0, gradient_path=capi.GRAD_AUTO, lib=None):
    """≙ mcmc_keep_warmup (documented, not exported, in the reference: src/mcmc.jl:23-50): MCMC with NUTS keeping the
    warmup results.  Returns a dict with
      initial_warmup_state, final_warmup_state : warmup states (q, κ, ϵ, generator position; ≙ WarmupState, src/warmup.jl:47-51)
      warmup    : list of dicts (stage, results = (chain, tree_statistics, ϵs) or None, warmup_state after the stage)
      inference : (chain [nchains, N, D], tree_statistics [nchains, N])
    Every warmup_state can be handed to `Engine.restore` / `initialization={"state": ...}` to continue from that point."""
    algorithm = algorithm or NUTS()
    initialization = initialization or {}
    if warmup_stages is None:
        warmup_stages = default_warmup_stages(stepsize_adaptation=DualAveraging(δ=δ))
    e = capi.Engine(nchains, ℓ.dim, dtype=dtype, max_depth=algorithm.max_depth, min_delta=algorithm.min_Δ, seed=seed,
                    chain_offset=chain_offset, device=device, gradient_path=gradient_path, lib=lib)
    ℓ.attach(e)
    if initialization.get("state") is not None:
        e.restore(initialization["state"])
    else:
        κ = initialization.get("κ")
        if κ is not None:
            e.set_metric_diag(np.broadcast_to(np.asarray(κ.M_inv, dtype=np.float64), (nchains, ℓ.dim)))
        e.set_positions(initialization.get("q"))
        if initialization.get("ϵ") is not None:
            e.set_stepsize(initialization["ϵ"])
    out = {"initial_warmup_state": e.warmup_state(), "warmup": []}
    for st in warmup_stages:
        if st is None:
            continue
        results = None
        if isinstance(st, FindLocalOptimum):
            e.find_local_optimum(st.magnitude_penalty, st.iterations)
        elif isinstance(st, InitialStepsizeSearch):
            e.find_initial_stepsize(st.a_min, st.a_max, st.ϵ0, st.C, st.maxiter_crossing, st.maxiter_bisect)
        elif isinstance(st, TuningNUTS):
            da = st.stepsize_adaptation
            fixed = isinstance(da, FixedStepsize)
            da = DualAveraging() if fixed else da
            results = e.warmup_stage(st.N, capi.METRIC_DIAG if st.M else capi.METRIC_NONE, da.δ, da.γ, da.κ, da.t0,
                                     -1.0 if st.λ is None else st.λ, fixed_stepsize=fixed)
        else:
            raise TypeError(f"unknown warmup stage {st!r}")
        out["warmup"].append({"stage": st, "results": results, "warmup_state": e.warmup_state()})
    out["final_warmup_state"] = e.warmup_state()
    out["inference"] = e.sample(N)
    e.close()
    return out


def mcmc_with_warmup(ℓ, N, **kw):
    """≙ mcmc_with_warmup(ℓ, N; ...), src/mcmc.jl:109-128: one chain; returns (chain [N, D], tree_statistics [N])."""
    kw.setdefault("nchains", 1)
    chains, stats = threaded_mcmc(ℓ, N, **kw)
    return chains[0], stats[0]

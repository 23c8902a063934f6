"""bnuts — B200-native batched-chain NUTS engine behind InplaceDHMC.jl's API names.

The directory name contains a dot, so import it through the root-level shim
``import inplacedhmc_jl_b200`` (see inplacedhmc_jl_b200.py).
"""
from ._capi import (Engine, BnutsError, load_library, nccl_unique_id, synth_logistic_rows, DEFAULT_LIB, TREE_STATS_DTYPE, EXPORTS,  # noqa: F401
                    F64, F32, X_F64, X_F32, X_BF16, GRAD_AUTO, GRAD_DETERMINISTIC, GRAD_TENSOR,
                    METRIC_NONE, METRIC_DIAG)
from .api import (threaded_mcmc, mcmc_with_warmup, mcmc_keep_warmup, fixed_stepsize_warmup_stages, FixedStepsize, default_warmup_stages, TuningNUTS, GaussianKineticEnergy,  # noqa: F401,E402
                  DualAveraging, InitialStepsizeSearch, FindLocalOptimum, NUTS, NoProgressReport, LogProgressReport,
                  IIDNormal, Funnel, Gaussian, Logistic)
from . import diagnostics  # noqa: F401,E402

"""Post-hoc diagnostics over TreeStatisticsNUTS records and draws.

EBFMI / summarize_tree_statistics / count_terminations / count_depths mirror
src/diagnostics.jl:28-101.  The effective-sample-size estimator has no counterpart in
the reference (it only offers E-BFMI); BASELINE.json's min-ESS/sec needs one, so the
standard multi-chain estimator is implemented here: rank-normalised draws, FFT
autocovariance, Geyer initial-monotone truncation, between/within-chain variance.
"""
from collections import namedtuple

import numpy as np

ACCEPTANCE_QUANTILES = [0.05, 0.25, 0.5, 0.75, 0.95]   # src/diagnostics.jl:35
MAX_DIRECTIONS_DEPTH = 32                               # src/tree.jl:132

TreeStatisticsSummary = namedtuple("TreeStatisticsSummary", "N a_mean a_quantiles termination_counts depth_counts")


def EBFMI(tree_statistics):
    """≙ EBFMI, src/diagnostics.jl:28-32: mean(abs2, diff(π)) / var(π); a matrix gives one value per chain."""
    pi = np.asarray(tree_statistics["pi"], dtype=np.float64)
    if pi.ndim == 1:
        return float(np.mean(np.diff(pi) ** 2) / np.var(pi, ddof=1))
    return np.array([EBFMI(tree_statistics[c]) for c in range(pi.shape[0])])


def count_terminations(tree_statistics):
    """≙ src/diagnostics.jl:61-76"""
    l = np.asarray(tree_statistics["term_left"]).ravel(); r = np.asarray(tree_statistics["term_right"]).ravel()
    max_depth = int(np.sum((l == 1) & (r == 0)))
    divergence = int(np.sum(l == r))
    return {"max_depth": max_depth, "divergence": divergence, "turning": int(l.size - max_depth - divergence)}


def count_depths(tree_statistics):
    """≙ src/diagnostics.jl:82-88"""
    c = np.bincount(np.asarray(tree_statistics["depth"]).ravel(), minlength=1)
    nz = np.nonzero(c)[0]
    return c[: (nz[-1] + 1) if nz.size else 0]


def summarize_tree_statistics(tree_statistics):
    """≙ src/diagnostics.jl:94-101"""
    a = np.asarray(tree_statistics["acceptance_rate"]).ravel()
    return TreeStatisticsSummary(a.size, float(a.mean()), np.quantile(a, ACCEPTANCE_QUANTILES),
                                 count_terminations(tree_statistics), count_depths(tree_statistics))


def _autocov_fft(x):
    n = x.shape[-1]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(axis=-1, keepdims=True)
    f = np.fft.rfft(xc, n=m, axis=-1)
    ac = np.fft.irfft(f * np.conj(f), n=m, axis=-1)[..., :n]
    return ac / n


def _rank_normalise(x):
    from scipy.special import ndtri
    flat = x.reshape(-1)
    ranks = np.empty_like(flat)
    order = np.argsort(flat, kind="mergesort")
    ranks[order] = np.arange(1, flat.size + 1)
    return ndtri((ranks - 0.375) / (flat.size + 0.25)).reshape(x.shape)


def ess(draws, rank_normalise=True):
    """Multi-chain bulk effective sample size of one scalar quantity; draws is [chains, n]."""
    x = np.asarray(draws, dtype=np.float64)
    if rank_normalise:
        x = _rank_normalise(x)
    m, n = x.shape
    acov = _autocov_fft(x)
    chain_var = acov[:, 0] * n / (n - 1.0)
    W = chain_var.mean()
    B_over_n = x.mean(axis=1).var(ddof=1) if m > 1 else 0.0
    var_plus = W * (n - 1.0) / n + B_over_n
    if not var_plus > 0:
        return float(m * n)
    rho = 1.0 - (W - acov.mean(axis=0)) / var_plus
    rho[0] = 1.0
    # Geyer: sums of adjacent pairs must be positive and non-increasing
    tau, prev, t = -1.0, np.inf, 0
    while t + 1 < n:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        pair = min(pair, prev)
        tau += 2.0 * pair
        prev = pair
        t += 2
    return float(m * n / max(tau, 1.0 / np.log10(max(m * n, 10))))


def min_ess(chains, **kw):
    """min over coordinates; chains is [nchains, N, D] (the layout bnuts_sample fills)."""
    c = np.asarray(chains)
    return min(ess(c[:, :, d], **kw) for d in range(c.shape[2]))

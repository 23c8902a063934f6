"""The algebra behind the remainder mode of the tensor-core logistic path (csrc/logistic_rm.cu), in numpy Float64:
Taylor coefficients of the residual about the reference, the decomposition of gradient and log density into the exact
D x D part + remainder, the identity that lets the consumer take a third of the log density's remainder from the gradient
partials, and the truncation error of the degree-4 form against the figures DESIGN.md section 6 states.  No GPU, no library."""
import numpy as np

from conftest import make_logistic


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def _coeffs(eta0):
    """(r0, w, A2, A3, A4) as k_rm_records computes them: u = 2 r0 - 1 = -tanh(eta0 / 2), w = (1 - u^2) / 4."""
    th = np.tanh(0.5 * eta0)
    u, w = -th, 0.25 * (1.0 - th * th)
    return 0.5 * (1.0 + u), w, -0.5 * w * u, -w * (u * u - 2.0 * w) / 6.0, -w * u * (u * u - 8.0 * w) / 24.0


def test_taylor_coefficients_are_the_derivatives_of_the_residual():
    eta0 = np.linspace(-6, 6, 49)
    r0, w, A2, A3, A4 = _coeffs(eta0)
    assert np.allclose(r0, _sig(-eta0), rtol=0, atol=1e-15) and np.allclose(w, _sig(eta0) * _sig(-eta0), atol=1e-15)
    # r(eta0 + d) - r0 + w d - d^2 (A2 + A3 d + A4 d^2) = O(d^5), with the sixth derivative of log sigma bounded by 1/2
    for d in (1e-2, 5e-2, 0.2):
        for s in (d, -d):
            rem = _sig(-(eta0 + s)) - r0 + w * s - s * s * (A2 + A3 * s + A4 * s * s)
            assert np.max(np.abs(rem)) < 0.25 / 120 * d ** 5 * 2.0 + 1e-16, (d, np.max(np.abs(rem)))
    # lambda = log sigma(eta0 + d) - f0 - r0 d + w d^2 / 2 = d^3 (A2/3 + A3 d/4 + A4 d^2/5) + O(d^6):  d lambda / d d = rho
    f = lambda x: np.minimum(x, 0) - np.log1p(np.exp(-np.abs(x)))
    d = 0.05
    lam = f(eta0 + d) - f(eta0) - r0 * d + 0.5 * w * d * d
    assert np.max(np.abs(lam - d ** 3 * (A2 / 3 + A3 * d / 4 + A4 * d * d / 5))) < 2e-11


def test_gradient_and_log_density_decompose_exactly():
    N, D, C = 4000, 12, 6
    X, y, beta = make_logistic(N, D)
    Xs = X * (2 * y - 1)[:, None]
    rng = np.random.default_rng(0)
    b0 = beta + rng.normal(size=D) * 0.02
    q = b0[None, :] + rng.normal(size=(C, D)) * 0.03
    eta0 = Xs @ b0
    r0, w, A2, A3, A4 = _coeffs(eta0)
    g0 = Xs.T @ r0
    H0 = (Xs * w[:, None]).T @ Xs
    l0 = (np.minimum(eta0, 0) - np.log1p(np.exp(-np.abs(eta0)))).sum()
    for c in range(C):
        db = q[c] - b0
        d = Xs @ db
        eta = eta0 + d
        g_true = Xs.T @ _sig(-eta)
        l_true = (np.minimum(eta, 0) - np.log1p(np.exp(-np.abs(eta)))).sum()
        rho = _sig(-eta) - r0 + w * d                              # the exact remainder (closed form of the far path)
        lam = (np.minimum(eta, 0) - np.log1p(np.exp(-np.abs(eta)))) - (np.minimum(eta0, 0) - np.log1p(np.exp(-np.abs(eta0)))) - r0 * d + 0.5 * w * d * d
        G_rem = Xs.T @ rho
        assert np.allclose(g0 - H0 @ db + G_rem, g_true, rtol=0, atol=1e-9)
        # what the kernel sums is mu = lambda - d rho / 3; the consumer restores the third from db . G_rem
        mu = lam - d * rho / 3.0
        assert abs(np.sum(d * rho) - db @ G_rem) < 1e-9
        assert abs(l0 + g0 @ db - 0.5 * db @ H0 @ db + (db @ G_rem) / 3.0 + mu.sum() - l_true) < 1e-8
        # Taylor form of the remainder at this distance (rms of d about 0.1): the figures of DESIGN.md section 6
        rho_t = d * d * (A2 + A3 * d + A4 * d * d)
        mu_t = -A3 * d ** 4 / 12.0
        s = np.sqrt(np.mean(d * d))
        rel = np.linalg.norm(Xs.T @ (rho_t - rho)) / np.linalg.norm(g_true - q[c] * 0 - 0)
        assert rel < 1e-2 * 15 * s ** 4 + 1e-12, (s, rel)
        assert abs(mu_t.sum() - mu.sum()) < 2.0 * np.sum(np.abs(A4) * np.abs(d) ** 5) + 1e-12


def test_one_bf16_term_of_the_remainder_is_enough_near_the_reference():
    """2^-9 relative rounding of rho, random over the rows: ~1e-6 of the gradient at one posterior sd for N / D = 1e3 here
    (~1e-7 at N = 1e6), three orders below what one bf16 term of r - r0 costs (the delta mode's 1.7e-3 sqrt(D/N))."""
    from conftest import to_bf16_grid
    N, D = 100_000, 100
    X, y, beta = make_logistic(N, D)
    Xs = X * (2 * y - 1)[:, None]
    b0 = beta.copy()
    for _ in range(6):
        s_ = _sig(X @ b0)
        H = (X * (s_ * (1 - s_))[:, None]).T @ X + np.eye(D)
        b0 = b0 + np.linalg.solve(H, X.T @ (y - s_) - b0)
    sd = 1.0 / np.sqrt(np.diag(H))
    db = np.random.default_rng(1).normal(size=D) * sd
    eta0 = Xs @ b0; d = Xs @ db
    r0, w, A2, A3, A4 = _coeffs(eta0)
    rho = _sig(-(eta0 + d)) - r0 + w * d
    g = Xs.T @ _sig(-(eta0 + d)) - (b0 + db)
    e_rho = np.linalg.norm(Xs.T @ (to_bf16_grid(rho) - rho)) / np.linalg.norm(g)
    dr = _sig(-(eta0 + d)) - r0
    e_delta = np.linalg.norm(Xs.T @ (to_bf16_grid(dr) - dr)) / np.linalg.norm(g)
    assert e_rho < 2e-6 and e_delta > 20 * e_rho, (e_rho, e_delta)   # measured 8.8e-7 and 5.3e-5 at this N / D (row-wise rms of d = 0.07)

"""An independent, literal Python transcription of the reference recursion
(adjacent_tree src/tree.jl:321-366, sample_trajectory :382-444, leaf src/NUTS.jl:176-191,
is_turning :148-170) on iid N(0,I) with unit metric, checked against the oracle with
injected momenta and directions (the reference's own test hooks, src/NUTS.jl:251-258)."""
import ctypes as C
import math

import numpy as np
import pytest


class Ref:
    def __init__(self, lib, seed, chain, t, eps, max_depth, min_delta=-1000.0):
        self.lib, self.seed, self.chain, self.t, self.eps, self.max_depth, self.min_delta = \
            lib, seed, chain, t, eps, max_depth, min_delta
        lib.bnuts_oracle_exponential.restype = C.c_double
        lib.bnuts_oracle_exponential.argtypes = [C.c_uint64] + [C.c_uint32] * 5

    def H(self, z):
        return -0.5 * z[0] @ z[0] - 0.5 * z[1] @ z[1]

    def leapfrog(self, z, eps):
        q, p = z
        pm = p - 0.5 * eps * q
        qn = q + eps * pm
        return (qn, pm - 0.5 * eps * qn)

    def logaddexp(self, x, y):
        if not (math.isfinite(x) and math.isfinite(y)):
            return x if x > y else y
        return float(np.logaddexp(x, y))

    def rand_bool(self, logprob, j, k, n):
        if logprob >= 0:
            return True
        return self.lib.bnuts_oracle_exponential(self.seed, self.chain, self.t, j, k, n) > -logprob

    def leaf(self, z, i, initial):
        d = 0.0 if initial else self.H(z) - self.pi0
        div = d < self.min_delta
        v = (-math.inf, 0) if initial else (min(d, 0.0), 1)
        return ((z, i), d, (z[1], z[1], z[1])), v, div

    def turning(self, tau):
        pm, pp, rho = tau
        return (rho @ pm < 0) | (rho @ pp < 0)

    def comb_tau(self, t1, t2, fwd):
        x, y = (t1, t2) if fwd else (t2, t1)
        return (x[0], y[1], x[2] + y[2])

    def adjacent(self, z, i, depth, fwd, j, base):
        ip = i + (1 if fwd else -1)
        if depth == 0:
            zn = self.leapfrog(z, self.eps if fwd else -self.eps)
            (zeta, w, tau), v, div = self.leaf(zn, ip, False)
            return (zeta, w, tau, zn, ip), v, (div, (ip, ip))
        tm, vm, (inv, it) = self.adjacent(z, i, depth - 1, fwd, j, base)
        if inv:
            return tm, vm, (inv, it)
        tp, vp, (inv, it) = self.adjacent(tm[3], tm[4], depth - 1, fwd, j, base + (1 << (depth - 1)))
        v = (self.logaddexp(vm[0], vp[0]), vm[1] + vp[1])
        if inv:
            return tp, v, (inv, it)
        tau = self.comb_tau(tm[2], tp[2], fwd)
        if self.turning(tau):
            return tp, v, (True, (ip, tp[4]))
        w = self.logaddexp(tm[1], tp[1])
        zeta = tp[0] if self.rand_bool(tp[1] - w, j, depth, base + (1 << depth)) else tm[0]
        return (zeta, w, tau, tp[3], tp[4]), v, (False, (1, 0))

    def sample(self, q, p, dirs):
        z = (q, p)
        self.pi0 = self.H(z)
        (zeta, w, tau), v, _ = self.leaf(z, 0, True)
        zm = zp = z
        im = ipl = 0
        depth, term = 0, (1, 0)
        while depth < self.max_depth:
            fwd = bool(dirs & 1); dirs >>= 1
            t, vn, (inv, it) = self.adjacent(zp if fwd else zm, ipl if fwd else im, depth, fwd, depth, 0)
            v = (self.logaddexp(v[0], vn[0]), v[1] + vn[1])
            if inv:
                term = it
                break
            if fwd:
                zp, ipl = t[3], t[4]
            else:
                zm, im = t[3], t[4]
            wn = self.logaddexp(w, t[1])
            if self.rand_bool(t[1] - w, depth, 0, 0):
                zeta = t[0]
            w = wn
            depth += 1
            tau = self.comb_tau(tau, t[2], fwd)
            if self.turning(tau):
                term = (im, ipl)
                break
        acc = min(1.0, math.exp(v[0]) / v[1])
        return zeta, self.H(zeta[0]), acc, term, depth, v[1]


@pytest.mark.parametrize("eps,max_depth", [(0.3, 5), (0.9, 4), (0.05, 3), (1.9, 6)])
def test_oracle_tree_matches_python_transcription(bn, oracle_lib, eps, max_depth):
    Cn, D, T = 12, 5, 6
    seed = 1234
    rng = np.random.default_rng(17)
    q0 = rng.normal(size=(Cn, D))
    p = rng.normal(size=(T, Cn, D))
    dirs = rng.integers(0, 2 ** 32, size=(T, Cn), dtype=np.uint64).astype(np.uint32)
    e = bn.Engine(Cn, D, max_depth=max_depth, lib=oracle_lib, seed=seed)
    e.model_iid_normal()
    e.set_positions(q0)
    e.set_stepsize(eps)
    e.inject(T, dirs, p)
    ch, st, sel = e.sample(T, want_index=True)
    kinds = set()
    for c in range(Cn):
        q = q0[c]
        for t in range(T):
            r = Ref(oracle_lib, seed, c, t, eps, max_depth)
            zeta, pi, acc, term, depth, steps = r.sample(q, p[t, c], int(dirs[t, c]))
            s = st[c, t]
            assert (s["term_left"], s["term_right"], s["depth"], s["steps"]) == (term[0], term[1], depth, steps)
            assert sel[c, t] == zeta[1]
            np.testing.assert_allclose(ch[c, t], zeta[0][0], atol=1e-12)
            assert s["pi"] == pytest.approx(pi, abs=1e-10)
            assert s["acceptance_rate"] == pytest.approx(acc, abs=1e-12)
            kinds.add("max" if term == (1, 0) else "div" if term[0] == term[1] else "turn")
            q = ch[c, t]
    assert "turn" in kinds or "max" in kinds


def test_worked_event_order_depth_and_indices(bn, oracle_lib):
    """dirs = 0b101 (fwd, bwd, fwd): a completed depth-3 tree spans positions -2..5."""
    D = 4
    e = bn.Engine(1, D, max_depth=3, lib=oracle_lib)
    e.model_iid_normal()
    e.set_positions(np.full((1, D), 0.1))
    e.set_stepsize(1e-3)           # tiny step: never turns, never diverges
    e.inject(1, np.array([[0b101]], dtype=np.uint32), np.ones((1, 1, D)))
    ch, st, sel = e.sample(1, want_index=True)
    s = st[0, 0]
    assert (s["depth"], s["steps"], s["term_left"], s["term_right"]) == (3, 7, 1, 0)   # REACHED_MAX_DEPTH
    assert -2 <= sel[0, 0] <= 5


def test_divergence_keeps_last_completed_tree(bn, oracle_lib):
    """Divergent leaf i: termination (i,i), steps counts it, depth not incremented (src/tree.jl:417)."""
    D = 3
    e = bn.Engine(1, D, max_depth=5, min_delta=-1e-3, lib=oracle_lib)   # absurdly strict threshold
    e.model_iid_normal()
    q0 = np.full((1, D), 0.5)
    e.set_positions(q0)
    e.set_stepsize(1.5)
    e.inject(1, np.array([[0b1]], dtype=np.uint32), np.full((1, 1, D), 2.0))
    ch, st, sel = e.sample(1, want_index=True)
    s = st[0, 0]
    assert s["term_left"] == s["term_right"] == 1 and s["depth"] == 0 and s["steps"] == 1
    assert sel[0, 0] == 0
    np.testing.assert_array_equal(ch[0, 0], q0[0])


from hypothesis import given, settings, strategies as hst, HealthCheck  # noqa: E402


@settings(max_examples=120, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(D=hst.integers(1, 9), max_depth=hst.integers(1, 6), eps=hst.floats(0.02, 2.5), min_delta=hst.sampled_from([-1000.0, -2.0, -0.2]),
       seed=hst.integers(0, 2 ** 31), rs=hst.integers(0, 2 ** 31))
def test_oracle_matches_transcription_on_random_configurations(bn, oracle_lib, D, max_depth, eps, min_delta, seed, rs):
    """The oracle against the independent literal Python transcription of the reference recursion (class Ref above)
    on random shapes, step sizes, depth caps, divergence thresholds, injected momenta and directions: termination
    code, depth, steps and the selected index exact; the draw to 1e-12."""
    Cn, T = 2, 3
    rng = np.random.default_rng(rs)
    q0 = rng.normal(size=(Cn, D))
    p = rng.normal(size=(T, Cn, D))
    dirs = rng.integers(0, 2 ** 32, size=(T, Cn), dtype=np.uint64).astype(np.uint32)
    e = bn.Engine(Cn, D, max_depth=max_depth, min_delta=min_delta, lib=oracle_lib, seed=seed)
    e.model_iid_normal()
    e.set_positions(q0)
    e.set_stepsize(eps)
    e.inject(T, dirs, p)
    ch, st, sel = e.sample(T, want_index=True)
    for c in range(Cn):
        q = q0[c]
        for t in range(T):
            r = Ref(oracle_lib, seed, c, t, eps, max_depth)
            r.min_delta = min_delta
            zeta, pi, acc, term, depth, steps = r.sample(q, p[t, c], int(dirs[t, c]))
            s = st[c, t]
            assert (s["term_left"], s["term_right"], s["depth"], s["steps"]) == (term[0], term[1], depth, steps)
            assert sel[c, t] == zeta[1]
            np.testing.assert_allclose(ch[c, t], zeta[0][0], atol=1e-12)
            q = ch[c, t]

"""All three random streams of a transition teacher-forced: directions, momenta and the merge exponentials
(≙ sample_tree(rng, ...; p, directions) with a scripted rng reaching randexp in rand_bool_logprob,
src/NUTS.jl:251-258, :32-34).  The scripted stream is consumed in the reference's CALL order — one value per merge with
logprob2 < 0, merges in the post-order of the recursion — which the independent Python transcription below does
literally (a list it pops from), while the product's state machine is iterative and the oracle keeps a counter."""
import numpy as np
import pytest

from test_tree_semantics import Ref


class ScriptedRef(Ref):
    """The transcription with a scripted rng: rand_bool pops the next exponential when (and only when) it needs one."""

    def __init__(self, *a, stream=(), **kw):
        super().__init__(*a, **kw)
        self.stream = list(stream)
        self.consumed = 0

    def rand_bool(self, logprob, j, k, n):
        if logprob >= 0:
            return True
        k0 = self.consumed
        self.consumed += 1
        if k0 < len(self.stream):
            return self.stream[k0] > -logprob
        return self.lib.bnuts_oracle_exponential(self.seed, self.chain, self.t, j, k, n) > -logprob


def _run(bn, lib, Cn, D, T, eps, max_depth, seed, q0, dirs, p, exps, dtype=0, **kw):
    e = bn.Engine(Cn, D, dtype=dtype, max_depth=max_depth, lib=lib, seed=seed, **kw)
    e.model_iid_normal()
    e.set_positions(q0)
    e.set_stepsize(eps)
    e.inject(T, dirs, p, exps)
    out = e.sample(T, want_index=True)
    e.close()
    return out


@pytest.mark.parametrize("libname", ["oracle", "hostemu"])
@pytest.mark.parametrize("eps,max_depth,n_exps", [(0.3, 5, 64), (0.9, 4, 32), (0.05, 4, 3), (0.6, 6, 1)])
def test_scripted_exponentials_match_transcription(bn, oracle_lib, hostemu_lib, libname, eps, max_depth, n_exps):
    lib = oracle_lib if libname == "oracle" else hostemu_lib
    Cn, D, T, seed = 7, 5, 5, 99
    rng = np.random.default_rng(23)
    q0 = rng.normal(size=(Cn, D))
    p = rng.normal(size=(T, Cn, D))
    dirs = rng.integers(0, 2 ** 32, size=(T, Cn), dtype=np.uint64).astype(np.uint32)
    exps = rng.exponential(size=(T, Cn, n_exps))
    ch, st, sel = _run(bn, lib, Cn, D, T, eps, max_depth, seed, q0, dirs, p, exps)
    used_fallback = False
    for c in range(Cn):
        q = q0[c]
        for t in range(T):
            r = ScriptedRef(oracle_lib, seed, c, t, eps, max_depth, stream=exps[t, c])
            zeta, pi, acc, term, depth, steps = r.sample(q, p[t, c], int(dirs[t, c]))
            used_fallback |= r.consumed > n_exps
            s = st[c, t]
            assert (s["term_left"], s["term_right"], s["depth"], s["steps"]) == (term[0], term[1], depth, steps)
            assert sel[c, t] == zeta[1], (c, t)
            np.testing.assert_allclose(ch[c, t], zeta[0][0], atol=1e-12)
            q = ch[c, t]
    if n_exps <= 3:
        assert used_fallback          # the case exists to cover "more draws consumed than were scripted"


def test_scripted_exponentials_change_the_selection(bn, oracle_lib):
    """Sanity: the injected stream is really used (huge exponentials always pick the new subtree's proposal,
    tiny ones never do, so the selected indices differ)."""
    Cn, D, T, seed, eps, depth = 16, 4, 4, 5, 0.4, 4
    rng = np.random.default_rng(3)
    q0 = rng.normal(size=(Cn, D)); p = rng.normal(size=(T, Cn, D))
    dirs = rng.integers(0, 2 ** 32, size=(T, Cn), dtype=np.uint64).astype(np.uint32)
    big = _run(bn, oracle_lib, Cn, D, 1, eps, depth, seed, q0, dirs[:1], p[:1], np.full((1, Cn, 40), 1e9))
    small = _run(bn, oracle_lib, Cn, D, 1, eps, depth, seed, q0, dirs[:1], p[:1], np.full((1, Cn, 40), 1e-30))
    assert (big[1]["steps"] == small[1]["steps"]).all()       # the trajectory itself does not depend on the draws
    assert (big[2] != small[2]).any()
    assert (small[2] == 0).all() or (np.abs(small[2]) <= np.abs(big[2])).any()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [0, 1])
def test_cuda_scripted_exponentials_bitwise(bn, oracle_lib, cuda_lib, dtype):
    """CUDA engine vs oracle with all three streams injected: draws, statistics and selected indices bit for bit."""
    Cn, D, T, seed, eps, depth = 33, 37, 6, 7, 0.35, 6
    rng = np.random.default_rng(29)
    q0 = rng.normal(size=(Cn, D)); p = rng.normal(size=(T, Cn, D))
    dirs = rng.integers(0, 2 ** 32, size=(T, Cn), dtype=np.uint64).astype(np.uint32)
    for n_exps in (80, 2):
        exps = rng.exponential(size=(T, Cn, n_exps))
        a = _run(bn, oracle_lib, Cn, D, T, eps, depth, seed, q0, dirs, p, exps, dtype=dtype)
        b = _run(bn, cuda_lib, Cn, D, T, eps, depth, seed, q0, dirs, p, exps, dtype=dtype)
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes()

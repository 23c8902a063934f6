"""Checkpoint / resume (≙ WarmupState, src/warmup.jl:47-51, and mcmc_keep_warmup,
src/mcmc.jl:23-50): the state (q, κ, ϵ) plus the position of the counter-based generator is everything a run needs —
a run continued in a fresh engine is bit-identical to an uninterrupted one."""
import numpy as np
import pytest

from conftest import set_model

F64, F32 = 0, 1


def _fresh(bn, lib, kind, dtype, C=5, D=11, **kw):
    e = bn.Engine(C, D, dtype=dtype, max_depth=6, lib=lib, seed=17, **kw)
    set_model(e, kind, D)
    return e


def _resume_protocol(bn, lib, kind, dtype, **kw):
    # uninterrupted: search, two warmup stages, 30 draws in three calls
    a = _fresh(bn, lib, kind, dtype, **kw)
    a.set_positions(None); a.find_initial_stepsize()
    a.warmup_stage(20, 0); a.warmup_stage(25, 1)
    full = [a.sample(12), a.sample(10), a.sample(8)]
    # interrupted twice: after the first warmup stage and in the middle of sampling
    b = _fresh(bn, lib, kind, dtype, **kw)
    b.set_positions(None); b.find_initial_stepsize()
    b.warmup_stage(20, 0)
    ck1 = b.warmup_state(); b.close()
    c = _fresh(bn, lib, kind, dtype, **kw); c.restore(ck1)
    c.warmup_stage(25, 1)
    part = [c.sample(12)]
    ck2 = c.warmup_state(); c.close()
    d = _fresh(bn, lib, kind, dtype, **kw); d.restore(ck2)
    part += [d.sample(10), d.sample(8)]
    return full, part, ck2, d


@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("kind", ["funnel", "gauss", "logit"])
@pytest.mark.parametrize("which", ["oracle", "hostemu"])
def test_resume_is_bitwise(bn, oracle_lib, hostemu_lib, which, kind, dtype):
    lib = oracle_lib if which == "oracle" else hostemu_lib
    full, part, ck, eng = _resume_protocol(bn, lib, kind, dtype)
    for (ch0, st0), (ch1, st1) in zip(full, part):
        assert ch0.tobytes() == ch1.tobytes() and st0.tobytes() == st1.tobytes()
    seed, t = eng.rng()
    assert seed == 17 and t == ck["next_transition"] + 18      # the generator position advanced by the 18 draws since
    bad = ck["q"].copy(); bad[2, 0] = np.nan
    with pytest.raises(bn.BnutsError):
        eng.restore({**ck, "q": bad})                           # a non-finite position is refused, as by set_positions


def test_mcmc_keep_warmup_states_are_resumable(bn, hostemu_lib):
    """mcmc_keep_warmup returns the state after every stage; restarting from the one before the last stage reproduces
    the rest of the run (last stage + inference) bit for bit."""
    stages = (bn.InitialStepsizeSearch(), bn.TuningNUTS(15, bn.DualAveraging()), bn.TuningNUTS(20, bn.DualAveraging(), M="Diagonal"),
              bn.TuningNUTS(10, bn.DualAveraging()))
    kw = dict(nchains=4, lib=hostemu_lib, seed=3)
    r = bn.mcmc_keep_warmup(bn.Funnel(6), 15, warmup_stages=stages, **kw)
    assert [w["stage"] for w in r["warmup"]] == list(stages) and r["warmup"][0]["results"] is None
    assert r["warmup"][1]["results"][0].shape == (4, 15, 6)
    s = bn.mcmc_keep_warmup(bn.Funnel(6), 15, warmup_stages=stages[-1:], initialization={"state": r["warmup"][2]["warmup_state"]}, **kw)
    for x, y in zip(r["warmup"][3]["results"], s["warmup"][0]["results"]):
        assert np.asarray(x).tobytes() == np.asarray(y).tobytes()
    assert r["inference"][0].tobytes() == s["inference"][0].tobytes() and r["inference"][1].tobytes() == s["inference"][1].tobytes()
    for k in ("q", "κ", "ϵ"):
        assert r["final_warmup_state"][k].tobytes() == s["final_warmup_state"][k].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("kind", ["funnel", "logit"])
def test_cuda_resume_is_bitwise(bn, cuda_lib, kind, dtype):
    full, part, _, _ = _resume_protocol(bn, cuda_lib, kind, dtype, gradient_path=1)
    for (ch0, st0), (ch1, st1) in zip(full, part):
        assert ch0.tobytes() == ch1.tobytes() and st0.tobytes() == st1.tobytes()

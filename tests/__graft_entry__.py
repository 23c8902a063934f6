"""Driver entry points: build() compiles everything for sm_100a (no GPU needed),
smoke() runs one small invocation of the hot path on cuda:0 and checks it against
the oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build():
    # the product: libbnuts.so (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo, see csrc/Makefile)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "inplacedhmc.jl_b200", "csrc")])
    # the checker: CPU oracle (C++ restatement) and the host emulation used by the CPU tests.
    # /root/reference is Julia with un-vendored dependencies: there is nothing compilable for oracle/_ref.
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "hostemu")])
    import inplacedhmc_jl_b200 as bn
    bn.load_library()          # fails loudly if the CUDA extension is missing
    return None


def smoke():
    import numpy as np
    import torch
    import inplacedhmc_jl_b200 as bn
    from conftest import make_logistic, ORACLE_SO
    assert torch.cuda.is_available()
    # 1) NUTS transitions on the funnel: device engine vs oracle, bit for bit
    outs = []
    for lib in (ORACLE_SO, None):
        e = bn.Engine(64, 20, dtype=bn.F64, max_depth=6, lib=lib, seed=1)
        e.model_funnel(); e.set_positions(None); e.set_stepsize(0.2)
        outs.append(e.sample(10, want_index=True))
    for a, b in zip(*outs):
        assert a.tobytes() == b.tobytes(), "device engine differs from oracle"
    # 2) tcgen05 logistic gradient + transitions vs the fp64 oracle gradient
    N, D, C = 4096, 100, 128
    X, y, beta = make_logistic(N, D)
    q = np.tile(beta, (C, 1)) * np.linspace(0.5, 1.5, C)[:, None]
    ref = bn.Engine(C, D, dtype=bn.F64, lib=ORACLE_SO); ref.model_logistic(X, y, 1.0); ref.set_positions(q)
    tc = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR); tc.model_logistic(X, y, 1.0); tc.set_positions(q)
    g0, g1 = ref.get_state()[1], tc.get_state()[1]
    rel = np.linalg.norm(g1 - g0, axis=1) / np.linalg.norm(g0, axis=1)
    assert rel.max() < 1e-5, rel.max()
    tc.set_stepsize(0.01)
    ch, st = tc.sample(3)
    assert np.isfinite(ch).all() and (st["steps"] > 0).all()
    print("smoke ok: funnel bit-exact; tensor-path gradient rel err %.2e; steps %d" % (rel.max(), st["steps"].sum()))
    return None


if __name__ == "__main__":
    {"build": build, "smoke": smoke}[sys.argv[1] if len(sys.argv) > 1 else "build"]()

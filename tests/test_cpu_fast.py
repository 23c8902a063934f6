"""The NON-parity vectorised CPU leg of bench.py (oracle/cpu_fast.cpp) against the bit-exact oracle, and the thread
plumbing of the CPU legs (torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm must not inherit it)."""
import ctypes
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT, make_logistic


def test_vectorised_gradient_matches_oracle(bn, oracle_lib):
    sys.path.insert(0, ROOT)
    import bench
    N, D, C = 3001, 37, 5                      # N not a multiple of the row block, D not a multiple of the vector width
    X, y, beta = make_logistic(N, D)
    rng = np.random.default_rng(5)
    q = beta[None, :] + rng.normal(size=(C, D)) * 0.3
    q[0] *= 40.0                               # saturated rows: exp(-|eta|) underflows towards 0
    e = bn.Engine(C, D, dtype=bn.F64, lib=oracle_lib); e.model_logistic(X, y, 1.0); e.set_positions(q)
    _, g0, l0 = e.get_state()
    lib = bench._fast_lib()
    Xs = (X * (2.0 * y - 1.0)[:, None]).astype(np.float32)   # bf16-grid values: exact in Float32
    g = np.zeros((C, D)); l = np.zeros(C)
    lib.cpufast_logistic_grad.restype = ctypes.c_int
    used = lib.cpufast_logistic_grad(ctypes.c_void_p(Xs.ctypes.data), ctypes.c_int64(N), D, ctypes.c_double(1.0),
                                     ctypes.c_void_p(q.ctypes.data), C, ctypes.c_void_p(g.ctypes.data), ctypes.c_void_p(l.ctypes.data))
    assert used >= 1
    assert np.max(np.linalg.norm(g - g0, axis=1) / np.linalg.norm(g0, axis=1)) < 1e-9
    assert np.max(np.abs(l - l0) / np.abs(l0)) < 1e-12


def test_vectorised_leapfrog_matches_oracle(bn, oracle_lib):
    sys.path.insert(0, ROOT)
    import bench
    N, D, C = 2000, 20, 3
    X, y, beta = make_logistic(N, D)
    rng = np.random.default_rng(6)
    q = beta[None, :] + rng.normal(size=(C, D)) * 0.05; p = rng.normal(size=(C, D))
    e = bn.Engine(C, D, dtype=bn.F64, lib=oracle_lib); e.model_logistic(X, y, 1.0); e.set_positions(q)
    g = e.get_state()[1].copy()
    q1, p1, g1, l1 = e.leapfrog(p, 2e-3, 5)
    lib = bench._fast_lib()
    Xs = (X * (2.0 * y - 1.0)[:, None]).astype(np.float32)
    qq, pp, l = q.copy(), p.copy(), np.zeros(C)
    lib.cpufast_logistic_leapfrog(Xs.ctypes.data, N, D, 1.0, 2e-3, 5, C, qq.ctypes.data, pp.ctypes.data, g.ctypes.data, l.ctypes.data)
    for a, b in ((qq, q1), (pp, p1), (g, g1)):
        assert np.max(np.abs(a - b)) / np.max(np.abs(b)) < 1e-9
    assert np.max(np.abs(l - l1) / np.abs(l1)) < 1e-12


def test_reference_arm_ignores_inherited_omp_num_threads():
    """bench.py --impl reference under OMP_NUM_THREADS=1 (what torch.distributed.run exports) still offers every host core
    to the oracle and says how many threads were busy."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    code = ("import sys; sys.path.insert(0, %r); import bench; lib = bench.force_omp_threads(bench.host_cores()); "
            "print(lib.cpufast_max_threads(), bench.host_cores())" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == out[1]


def test_bench_flags_named_in_the_docs_exist():
    """DESIGN.md / profiles/README.md quote bench.py command lines: every flag they name is one the parser knows."""
    import re
    import subprocess
    import sys
    help_text = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, check=True).stdout
    known = set(re.findall(r"--[a-z][a-z0-9-]+", help_text))
    named = set()
    for doc in ("DESIGN.md", "profiles/README.md", "README.md", "BASELINE.md"):
        with open(os.path.join(ROOT, doc)) as f:
            for line in f:
                for cmd in re.findall(r"bench\.py((?: --?[A-Za-z0-9|\\=-]+(?: [A-Za-z0-9_.]+)?)+)", line):
                    named |= set(re.findall(r"--[a-z][a-z0-9-]+", cmd))
    assert named, "no bench.py command line found in the docs"
    assert named <= known, sorted(named - known)

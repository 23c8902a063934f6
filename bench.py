#!/usr/bin/env python
"""bench.py — leapfrog steps/s of the NUTS hot path on the BASELINE configurations.

  python bench.py --gpus N --steps K --warmup W              our CUDA engine, config c3 (the headline)
  python bench.py --config c2|c4|c5 ...                      the other BASELINE configurations (same JSON schema)
  python bench.py --impl reference ...                       CPU restatement of the reference on the host cores

c3 (default): Bayesian logistic regression N=1e6, D=100, 4096 chains sharded over the GPUs, no collective.
c2: 1000-dim correlated Gaussian, shared dense metric, 4096 chains (sharded).   c4: Neal's funnel D=100, 8192 chains
(sharded).   c5: logistic regression N=1e8, D=256, rows sharded over the GPUs, one exchange per leapfrog, all chains
replicated on every GPU.

A "step" is `--transitions` NUTS transitions of every chain (one bnuts_sample call): momentum refresh, tree building
(leapfrog + gradient per leaf), selection, statistics.  Chains run their transitions asynchronously inside the call
(a chain starts its next tree as soon as it finishes one); idle chains occupy no rows of the gradient kernel.
`value` = leapfrog steps (Σ TreeStatisticsNUTS.steps, src/NUTS.jl:238-239) of all chains ÷ device time of the K timed
steps, state resident in HBM.  `e2e` = the same through the C ABI with host buffers: positions are uploaded and draws
+ statistics downloaded inside the timed region.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

UNIT = "leapfrog steps/s"
METRICS = {
    "c3": "leapfrog steps/sec (4096 chains, logistic reg N=1e6 D=100)",
    "c2": "leapfrog steps/sec (4096 chains, correlated Gaussian D=1000, dense metric)",
    "c4": "leapfrog steps/sec (8192 chains, Neal's funnel D=100)",
    "c5": "leapfrog steps/sec (4096 chains, logistic reg N=1e8 D=256, rows sharded)",
}


def synth(N, D, seed=3):
    """X ~ N(0,1) on the bf16 grid, column 0 = 1, beta* ~ N(0,1/D), y ~ Bernoulli(sigmoid(X beta*))."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D), dtype=np.float32)
    u = X.view(np.uint32)
    bits = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)       # bf16 RNE
    bits[:, 0] = 0x3F80                                                    # 1.0
    Xf = (bits.astype(np.uint32) << 16).view(np.float32)
    beta = rng.standard_normal(D) / np.sqrt(D)
    eta = Xf.astype(np.float64) @ beta
    y = (rng.uniform(size=N) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
    return bits, y, beta


def synth_gaussian(D, seed=2):
    """A ~ N(0,1)^{D x 2D}, Sigma = A A'/2D, model precision P = Sigma^-1."""
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(D, 2 * D))
    S = A @ A.T / (2 * D)
    return np.linalg.inv(S), S


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU legs
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _fast_lib():
    from conftest import build_oracle
    build_oracle()
    path = os.path.join(ROOT, "oracle", "libbnuts_cpufast.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(path)
    lib.cpufast_max_threads.restype = ctypes.c_int
    lib.cpufast_logistic_leapfrog.restype = ctypes.c_int
    lib.cpufast_logistic_leapfrog.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                              ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
    return lib


def force_omp_threads(cores):
    """torchrun exports OMP_NUM_THREADS=1 to its workers, and libgomp may already be initialised (torch) when the
    oracle is loaded: set the environment unconditionally AND the run-time ICV (omp_set_num_threads of the calling
    thread, through the libgomp the oracle shares), then read it back."""
    os.environ["OMP_NUM_THREADS"] = str(cores)
    lib = _fast_lib()
    lib.cpufast_set_threads(int(cores))
    got = int(lib.cpufast_max_threads())
    assert got == cores, f"OpenMP offers {got} threads, wanted {cores}"
    return lib


def _timed(fn):
    """wall seconds and the number of host threads that were busy (process CPU time ÷ wall time)"""
    c0 = os.times(); t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0; c1 = os.times()
    return dt, ((c1.user - c0.user) + (c1.system - c0.system)) / max(dt, 1e-9)


def cpu_oracle_leapfrogs(model_fn, D, budget_s, dtype=0, seed=1, qscale=0.01, eps=1e-3, max_steps=64, cores=None):
    """The oracle (CPU restatement of src/kinetic_energy.jl:126-163 + the target), one chain per host thread like
    Threads.@threads in src/mcmc.jl:150-157, on a bounded number of bare leapfrog steps.  Returns
    (chain-leapfrogs/s, cores, busy threads measured, description of the sample)."""
    import inplacedhmc_jl_b200 as bn
    from conftest import build_oracle
    cores = cores or host_cores()
    force_omp_threads(cores)
    lib = bn.load_library(build_oracle())
    e = bn.Engine(cores, D, dtype=dtype, lib=lib, seed=seed)
    model_fn(e)
    rng = np.random.default_rng(seed)
    e.set_positions(rng.normal(size=(cores, D)) * qscale)
    p = rng.normal(size=(cores, D))
    t1, _ = _timed(lambda: e.leapfrog(p, eps, 1))
    n = int(max(1, min(max_steps, budget_s / max(t1, 1e-4))))
    dt, busy = _timed(lambda: e.leapfrog(p, eps, n))
    e.close()
    return cores * n / dt, cores, busy, f"{n} bare leapfrog steps x {cores} chains (one per host thread)"


def cpu_oracle_nuts(model_fn, D, chains, transitions, dtype=0, seed=1, max_depth=10, warm=((30, 0), (40, 1), (30, 0)), cores=None):
    """NUTS transitions on the oracle (cheap targets: c2, c4): leapfrog steps/s over `transitions` draws of `chains` chains."""
    import inplacedhmc_jl_b200 as bn
    from conftest import build_oracle
    cores = cores or host_cores()
    force_omp_threads(cores)
    lib = bn.load_library(build_oracle())
    e = bn.Engine(chains, D, dtype=dtype, lib=lib, seed=seed, max_depth=max_depth)
    model_fn(e)
    e.set_positions(None)
    e.find_initial_stepsize()
    for n, mk in warm:
        e.warmup_stage(n, mk, keep=False)
    c0 = e.counters()
    dt, busy = _timed(lambda: e.sample(transitions))
    leap = e.counters()["leapfrogs"] - c0["leapfrogs"]
    e.close()
    return leap / dt, cores, busy, f"{transitions} NUTS transitions x {chains} chains (one chain per host thread at a time), {leap} leapfrogs"


def cpu_vectorised_logistic(bits, y, D, budget_s, seed=1, cores=None):
    """NON-parity leg (oracle/cpu_fast.cpp): the same leapfrog with a SIMD gradient, one chain per host thread."""
    cores = cores or host_cores()
    lib = force_omp_threads(cores)
    Xs = ((bits.astype(np.uint32) << 16).view(np.float32) * (2.0 * y - 1.0).astype(np.float32)[:, None]).astype(np.float32)
    N = Xs.shape[0]
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(cores, D)) * 0.01; p = rng.normal(size=(cores, D)); g = np.zeros((cores, D)); l = np.zeros(cores)

    def leap(n):
        return lib.cpufast_logistic_leapfrog(Xs.ctypes.data, N, D, 1.0, 1e-3, n, cores, q.ctypes.data, p.ctypes.data, g.ctypes.data, l.ctypes.data)
    t1, _ = _timed(lambda: leap(1))
    n = int(max(1, min(256, budget_s / max(t1, 1e-4))))
    used = [0]
    dt, busy = _timed(lambda: used.__setitem__(0, leap(n)))
    return {"value": cores * n / dt, "unit": UNIT, "cores": cores, "threads_used": int(used[0]), "busy_threads_measured": round(busy, 1),
            "kind": "vectorised, not a parity reference (oracle/cpu_fast.cpp: Float64 SIMD gradient, no fixed summation order)",
            "sample": f"{n} bare leapfrog steps x {cores} chains (one per host thread), N={N} D={D}"}


def logistic_model_fn(bits, y):
    return lambda e: e.model_logistic(bits, y, 1.0, row_blocks=1)


def reference_leg(a, budget_s):
    """CPU arm of the configuration: (value, cores, busy, sample text, extra dict)."""
    cfg = a.config
    if cfg == "c3":
        bits, y, _ = synth(a.rows, a.dim)
        v, cores, busy, sample = cpu_oracle_leapfrogs(logistic_model_fn(bits, y), a.dim, budget_s)
        extra = {}
        if not a.no_vectorised:
            extra["vectorised"] = cpu_vectorised_logistic(bits, y, a.dim, min(budget_s, 10.0))
        return v, cores, busy, sample + f", N={a.rows} D={a.dim}, fp64 oracle", extra
    if cfg == "c5":
        # bounded sample: a slice of the rows (the gradient is a sum over rows, its cost is linear in them); scaled to N
        import inplacedhmc_jl_b200 as bn
        ns = 200_000
        v, cores, busy, sample = cpu_oracle_leapfrogs(lambda e: e.model_logistic_synthetic(5, 0, ns, 1.0), a.dim, budget_s)
        return v * ns / a.rows, cores, busy, sample + f" on rows [0, {ns}) of the synthetic matrix, D={a.dim}; rate scaled by {ns}/{a.rows} rows", {}
    if cfg == "c2":
        P, S = synth_gaussian(a.dim)
        def fn(e):
            e.model_gaussian(P); e.set_metric_dense(S)
        cores = host_cores()
        v, cores, busy, sample = cpu_oracle_nuts(fn, a.dim, cores, 8, warm=((30, 0),), cores=cores)
        return v, cores, busy, sample + f", D={a.dim} dense Gaussian, dense metric, fp64 oracle", {}
    if cfg == "c4":
        cores = host_cores()
        v, cores, busy, sample = cpu_oracle_nuts(lambda e: e.model_funnel(), a.dim, 4 * cores, 200, cores=cores)
        return v, cores, busy, sample + f", funnel D={a.dim}, fp64 oracle", {}
    raise ValueError(cfg)


def workload_name(a, chains_total):
    return {
        "c3": "c3: Bayesian logistic regression N=%d D=%d, %d chains total, max_depth 10" % (a.rows, a.dim, chains_total),
        "c2": "c2: %d-dim correlated Gaussian (dense Sigma), shared dense-metric GaussianKE, %d chains total, max_depth 10" % (a.dim, chains_total),
        "c4": "c4: Neal's funnel %d-dim, %d chains total, max_depth 10" % (a.dim, chains_total),
        "c5": "c5: logistic regression N=%d D=%d, %d chains, rows sharded over the GPUs with a per-leapfrog gradient exchange" % (a.rows, a.dim, chains_total),
    }[a.config]


def run_reference(a, rank, world):
    if rank != 0:
        return
    rates, sample, cores, busy, extra = [], "", 1, 0.0, {}
    t0 = time.perf_counter()
    for i in range(a.warmup + a.steps):
        r, cores, busy, sample, extra = reference_leg(a, budget_s=max(2.0, 60.0 / (a.warmup + a.steps)))
        if i >= a.warmup:
            rates.append(r)
    val = float(np.mean(rates))
    ms = (time.perf_counter() - t0) * 1e3 / (a.warmup + a.steps)
    cb = {"value": val, "unit": UNIT, "cores": cores, "busy_threads_measured": round(busy, 1), "kind": "port", "sample": sample}
    cb.update(extra)
    print(json.dumps({
        "impl": "reference", "metric": METRICS[a.config], "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a, a.chains),
                   "sample": "bounded sample of the workload per step on the host cores (see cpu_baseline.sample); "
                             "bit-exact scalar port of the reference (-O2 -mfma -mavx2 -ffp-contract=off), its vectorised twin beside it"},
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def setup_engine(a, bn, rank, world, local):
    """Build the engine of the configuration and run its (untimed) warmup.  Returns (engine, chains on this rank, info)."""
    cfg = a.config
    info = {}
    t_w = time.perf_counter()
    if cfg in ("c3", "c5"):
        D = a.dim
        if cfg == "c3" and a.sharding == "rows" and world > 1:
            # OPTION (not BASELINE config 3's layout, which shards chains): config 5's layout on config 3's problem — every GPU holds
            # all chains and 1/world of the rows (30 MB at 8 GPUs: L2-resident), one NCCL exchange per leapfrog
            C = a.chains
            bits, y, beta = synth(a.rows, D)
            lo, hi = a.rows * rank // world, a.rows * (rank + 1) // world
            e = bn.Engine(C, D, dtype=bn.F32, seed=20261018, device=local, gradient_path=bn.GRAD_TENSOR)
            e.model_logistic(bits[lo:hi], y[lo:hi], 1.0)
            info["cpu_inputs"] = (bits, y)
            info["rows_this_rank"] = hi - lo
            import torch.distributed as dist
            ids = [bn.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            e.set_nccl(ids[0], world, rank)
        elif cfg == "c3":
            C = a.chains // world if a.scaling == "strong" else a.chains   # chains shard across GPUs, no communication
            bits, y, beta = synth(a.rows, D)
            e = bn.Engine(C, D, dtype=bn.F32, seed=20261018, chain_offset=rank * C, device=local, gradient_path=bn.GRAD_TENSOR)
            e.model_logistic(bits, y, 1.0)
            info["cpu_inputs"] = (bits, y)
        else:
            C = a.chains                                                   # all chains on every GPU, rows sharded
            r0 = a.rows * rank // world; r1 = a.rows * (rank + 1) // world
            e = bn.Engine(C, D, dtype=bn.F32, seed=20261018, device=local, gradient_path=bn.GRAD_TENSOR)
            e.model_logistic_synthetic(5, r0, r1 - r0, 1.0)
            info["rows_this_rank"] = r1 - r0
            if world > 1:
                import torch.distributed as dist
                if a.exchange == "nccl":
                    ids = [bn.nccl_unique_id() if rank == 0 else None]
                    dist.broadcast_object_list(ids, src=0)
                    e.set_nccl(ids[0], world, rank)
                else:
                    hs = [None] * world
                    dist.all_gather_object(hs, e.p2p_export())
                    e.p2p_connect(hs, rank)
        # ≙ default_warmup_stages (src/warmup.jl:361-372): q0 ~ U[-2,2]^D (:73), FindLocalOptimum, InitialStepsizeSearch, ...
        e.set_positions(None)
        e.find_local_optimum(1e-4, a.opt_iters)
        # reference point of the tensor path (include/bnuts.h) = the optimum just found (across-chain mean); the engine
        # checks it and keeps the exact three-term path if it is refused.  D > 128 always works about a reference (16-bit
        # operand about zero at first), so there the search is repeated from the more accurate operand
        info["position_operand_terms"] = 3 if D <= 128 else 2
        if not a.no_reference:
            try:
                e.logistic_set_reference(e.get_state()[0].mean(axis=0))
                info["position_operand_terms"] = 2
                for _ in range(2 if D > 128 else 0):
                    e.find_local_optimum(1e-4, a.opt_iters)
                    e.logistic_set_reference(e.get_state()[0].mean(axis=0))
            except bn.BnutsError as ex:
                print("reference point refused: %s" % ex, file=sys.stderr)
        e.find_initial_stepsize()
        # ≙ default_warmup_stages with shorter windows: step size only, then step size + per-chain diagonal metric in
        # doubling windows, then step size only
        stages = ((a.adapt, bn.METRIC_NONE), (25, bn.METRIC_DIAG), (50, bn.METRIC_DIAG), (100, bn.METRIC_DIAG), (a.adapt, bn.METRIC_NONE)) \
            if cfg == "c3" else ((a.adapt, bn.METRIC_NONE),)
        if cfg == "c3" and a.full_warmup:   # the windows of default_warmup_stages itself (src/warmup.jl:361-372): 75 | 25 50 100 200 400 | 50
            stages = ((75, bn.METRIC_NONE),) + tuple((n, bn.METRIC_DIAG) for n in (25, 50, 100, 200, 400)) + ((50, bn.METRIC_NONE),)
        info["init"] = "q0 ~ U[-2,2]^D; untimed warmup: FindLocalOptimum(1e-4, %d), step size search, stages %s" % (
            a.opt_iters, "|".join("%d%s" % (n, "m" if mk else "") for n, mk in stages))
    elif cfg == "c2":
        D = a.dim
        C = a.chains // world
        P, S = synth_gaussian(D)
        e = bn.Engine(C, D, dtype=bn.F32, seed=20261018, chain_offset=rank * C, device=local, gradient_path=bn.GRAD_TENSOR)
        e.model_gaussian(P)
        e.set_metric_dense(S)                                              # c2: M^-1 = Sigma injected, shared
        e.set_positions(None)
        e.find_initial_stepsize()
        stages = ((75, bn.METRIC_NONE), (100, bn.METRIC_NONE))
        info["init"] = "q0 ~ U[-2,2]^D; M^-1 = Sigma (dense, shared); untimed warmup: step size search, 75|100 step-size-only stages"
    elif cfg == "c4":
        D = a.dim
        C = a.chains // world
        e = bn.Engine(C, D, dtype=bn.F64 if a.dtype == "f64" else bn.F32, seed=20261018, chain_offset=rank * C, device=local)
        e.model_funnel()
        e.set_positions(None)
        e.find_initial_stepsize()
        stages = ((75, bn.METRIC_NONE), (25, bn.METRIC_DIAG), (50, bn.METRIC_DIAG), (100, bn.METRIC_DIAG), (50, bn.METRIC_NONE))
        info["init"] = "q0 ~ U[-2,2]^D; untimed warmup: step size search, 75|25m|50m|100m|50 (delta 0.9)"
    else:
        raise ValueError(cfg)
    if a.adapt > 0 or cfg in ("c2", "c4"):
        for n, mk in stages:
            if n > 0:
                e.warmup_stage(n, mk, keep=False, delta=0.9 if cfg == "c4" else 0.8)
    info["warmup_s"] = time.perf_counter() - t_w
    return e, C, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--chains", type=int, default=None)
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--adapt", type=int, default=None, help="length of the first/last step-size-only warmup stage (untimed)")
    ap.add_argument("--opt-iters", type=int, default=None, help="iterations of FindLocalOptimum (untimed)")
    ap.add_argument("--transitions", type=int, default=None, help="NUTS transitions per chain per step")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="c4 only: engine arithmetic")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"], help="c5 only: per-leapfrog exchange of the gradient partials")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vectorised", action="store_true", help="skip the non-parity vectorised CPU leg")
    ap.add_argument("--no-fp64", action="store_true", help="c3: skip the Float64-engine leg (reference precision)")
    ap.add_argument("--no-reference", action="store_true", help="keep the exact three-term position operand")
    ap.add_argument("--sharding", default="chains", choices=["chains", "rows"],
                    help="c3 on several GPUs: chains (BASELINE config 3, default) or rows with a per-leapfrog NCCL exchange (an option: every GPU advances all chains)")
    ap.add_argument("--full-warmup", action="store_true", help="c3: the windows of default_warmup_stages (75|25,50,100,200,400|50) instead of the shortened 40|25,50,100|40")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (BASELINE config 3: --chains in total, sharded over the GPUs) or weak (--chains per GPU)")
    a = ap.parse_args()
    dflt = {"c3": dict(chains=4096, rows=1_000_000, dim=100, adapt=40, opt_iters=50, transitions=64),
            "c2": dict(chains=4096, rows=0, dim=1000, adapt=0, opt_iters=0, transitions=64),
            "c4": dict(chains=8192, rows=0, dim=100, adapt=0, opt_iters=0, transitions=100),
            "c5": dict(chains=4096, rows=100_000_000, dim=256, adapt=4, opt_iters=30, transitions=1)}[a.config]
    for k, v in dflt.items():
        if getattr(a, k) is None:
            setattr(a, k, v)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if a.warmup < 3:
        a.warmup = 3

    import torch
    import inplacedhmc_jl_b200 as bn
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t[0])

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t[0])

    replicated = a.config == "c5" or (a.config == "c3" and a.sharding == "rows" and world > 1)   # every rank runs the same chains: count them once
    e, C, info = setup_engine(a, bn, rank, world, local)
    D, T = a.dim, a.transitions
    for _ in range(a.warmup):
        e.sample_device_only(T)

    # ---------------- timed region 1: device-resident
    clocks = ClockSampler(local); clocks.start()
    c0 = e.counters(); e.profile(True)
    barrier()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        e.sample_device_only(T)
    ev1.record(); barrier()
    ms = allmax(ev0.elapsed_time(ev1))
    grad_ms, grad_n = e.profile(False)
    c1 = e.counters()
    clk = clocks.stop()
    leap_local = c1["leapfrogs"] - c0["leapfrogs"]
    leap = leap_local if replicated else allsum(leap_local)
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    value = leap / (ms * 1e-3)

    # ---------------- timed region 2: through the C ABI with host buffers
    qh = torch.empty((C, D), dtype=torch.float64).pin_memory().numpy()
    qh[:] = e.get_state()[0]
    chain = torch.empty((C, T, D), dtype=torch.float64).pin_memory().numpy()
    stats = np.zeros((C, T), dtype=bn.TREE_STATS_DTYPE)
    keep_draws = a.config == "c3"
    kept = np.empty((C, a.steps * T, D)) if keep_draws else None   # the e2e draws form one continuous chain per chain id
    barrier()
    t0 = time.perf_counter(); leap_e2e = 0
    for k in range(a.steps):
        e.set_positions(qh)                       # H2D of this step's input positions (+ their gradient)
        e.sample(T, out=(chain, stats))           # D2H of the draws and tree statistics
        leap_e2e += int(stats["steps"].sum())
        qh[:] = chain[:, T - 1]
        if keep_draws:
            kept[:, k * T:(k + 1) * T] = chain
    barrier()
    dt = allmax(time.perf_counter() - t0)
    e2e = (leap_e2e if replicated else allsum(leap_e2e)) / dt
    replicas_equal = None
    if replicated and dist is not None:   # every rank ran the same chains on all-reduced gradients: the outputs must be the same bits
        import hashlib
        hs = [None] * world
        dist.all_gather_object(hs, hashlib.sha256(chain.tobytes() + stats.tobytes()).hexdigest())
        replicas_equal = len(set(hs)) == 1
    min_ess = None
    if keep_draws:
        # min-ESS/s (second half of BASELINE.json's metric): multi-chain bulk ESS per coordinate of the e2e draws;
        # chains on different GPUs are independent, so per-coordinate ESS adds across ranks
        # (estimated on the first 512 chains of the rank and scaled to all of them: chains are i.i.d. replicas)
        nsub = min(C, 512)
        ess_d = (np.array([bn.diagnostics.ess(kept[:nsub, :, d]) for d in range(D)]) * (C / nsub)) if kept.shape[1] >= 4 else np.zeros(D)
        if dist is not None and not replicated:
            tt = torch.tensor(ess_d, dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            ess_d = tt.cpu().numpy()
        min_ess = float(ess_d.min())

    # ---------------- c3: the same model through the Float64 engine (the reference is Float64-only, src/warmup.jl:108-120)
    fp64 = None
    if a.config == "c3" and rank == 0 and not a.no_fp64:
        bits, y = info["cpu_inputs"]
        C64 = min(C, 1024)
        e64 = bn.Engine(C64, D, dtype=bn.F64, seed=20261018, device=local, gradient_path=bn.GRAD_DETERMINISTIC)
        e64.model_logistic(bits, y, 1.0, row_blocks=64)
        e64.set_positions(e.get_state()[0][:C64])
        p = np.random.default_rng(7).normal(size=(C64, D))
        e64.leapfrog(p, 1e-3, 1)
        torch.cuda.synchronize(); t1 = time.perf_counter(); e64.leapfrog(p, 1e-3, 4); torch.cuda.synchronize()
        d64 = time.perf_counter() - t1
        fp64 = {"value": C64 * 4 / d64, "unit": UNIT,
                "sample": "4 bare leapfrog steps x %d chains, Float64 engine, deterministic CUDA-core gradient (k_grad_logistic<double>, bit-exact with the oracle)" % C64}
        e64.close()

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    lock = int(c1["lockstep_steps"] - c0["lockstep_steps"])
    rows = c1["gradient_rows"] - c0["gradient_rows"]   # active chains summed over the gradient launches
    chains_total = C if replicated else C * world
    config = {"workload": workload_name(a, chains_total), "chains_per_gpu": C,
              "parallelism": ("rows sharded over %d GPUs, chains replicated, per-leapfrog exchange (%s)" % (world, a.exchange if a.config == "c5" else "nccl")) if replicated
              else "chains sharded, no collective",
              "init": info["init"] + ", %.1f s" % info["warmup_s"],
              "step": "%d NUTS transitions of every chain (async within the call)" % T,
              "mean_tree_depth": float(stats["depth"].mean()), "mean_leapfrogs_per_transition": float(stats["steps"].mean()),
              "lockstep_steps_timed": lock,
              "active_row_fraction": float(leap_local / max(1, lock * C)) if a.config != "c4" else None}
    if a.config == "c3":
        # measured (profiles/r2_tail_probe_summary.json): a chain's mean tree length is persistent (NUTS resonance
        # of a near-isotropic posterior), so the lockstep count of a call is the sequential depth of its slowest chain whatever the
        # scheduling; one call over all transitions would give 0.225 (0.108 with the reference's own warmup windows, --full-warmup,
        # which leave a MORE ragged workload: 1.10 M leapfrog steps/s in profiles/r2_bench_c3_1gpu_full_warmup.json)
        config["active_row_fraction_bound"] = "persistent per-chain tree lengths: <= 0.23 for this workload under any scheduling (DESIGN.md section 5)"
        config["warmup_windows"] = "reference default 75|25,50,100,200,400|50 (--full-warmup)" if a.full_warmup else "shortened 40|25,50,100|40 (the reference's default windows: --full-warmup)"
    out = {
        "metric": METRICS[a.config], "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "data": "synthetic", "config": config, "gpu_launches": int(launches), "leapfrogs_timed": int(leap),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(qh.nbytes), "d2h_bytes_per_step": int(chain.nbytes + stats.nbytes)},
        "clocks": clk,
    }
    if a.config in ("c3", "c5"):
        N = a.rows
        N_gpu = info.get("rows_this_rank", a.rows)
        peak_tf = peaks.get("bf16_tflops_sustained")
        peak_src = "measured sustained (MEASURED_PEAKS.json)" if peak_tf else "fallback 1400 (B200_PROFILING.md)"
        peak_tf = peak_tf or 1400.0
        terms = info["position_operand_terms"]
        kern = "k_logistic_tc" if D <= 128 else "k_logistic_tc256"
        # algorithmic flops per GPU: X·B and Xᵀ·R with the true D, 2 flop per MAC, only rows that were requested
        ach = 4.0 * N_gpu * D * rows / (grad_ms * 1e-3) / 1e12 if grad_n else None
        dk = (-(-(D + 3) // 16) * 16) if D <= 125 else (-(-D // 16) * 16)
        rterms = 2 if (a.config == "c3") else 1
        # remainder mode (csrc/logistic_rm.cu): what logistic_tc_set_reference takes by itself for D <= 125, N >= 3000 D, rows not sharded
        rmode_env = os.environ.get("BNUTS_TC_RMODE")
        remainder = (terms == 2 and (D <= 125 or D > 128) and (a.config == "c3" or a.exchange == "nccl" or world == 1)
                     and (int(rmode_env) == 2 if rmode_env else (N >= 3000 * D and "BNUTS_TC_RREF" not in os.environ)))
        if remainder:
            kern = "k_logistic_rm"
        # executed on the tensor pipe per algorithmic flop: K padded to 16 and the position operand in `terms` bf16 terms in GEMM1;
        # the residual in `rterms` terms with N = dk (k_logistic_tc) or the remainder in one term with M = 128 features (k_logistic_rm) in GEMM2
        exe = ((terms * dk + (128 if D <= 128 else 256)) / (2.0 * D)) if remainder else ((terms + rterms) * dk / (2.0 * D))
        out["dtype"] = "f32 (exact bf16 operand splits on tcgen05, fp32 accumulate)"
        config["l2"] = "inputs larger than L2 (X is %d MB bf16 per GPU)" % (N_gpu * (128 if D <= 128 else 256) * 2 // 2**20)
        config["position_operand_terms"] = terms
        out["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                           "frac": (ach / peak_tf) if ach else None,
                           # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full captures
                           # (k_logistic_rm: profiles/r2_ncu_full_details_k_logistic_rm_4096rows.csv, 306.2 MB + 14.7 MB against 256 MB of
                           # X + 16 MB of per-row records algorithmic; k_logistic_tc: profiles/r1_..._k_logistic_tc_4096rows.csv,
                           # 292.4 MB + 14.5 MB); X is read once per launch whatever the number of active rows
                           "traffic": (320.8e6 if remainder else 306.9e6) if (a.config == "c3" and N == 1_000_000 and D == 100) else None,
                           "traffic_unit": "bytes per launch (ncu, full 4096-row launch)",
                           "kernel": kern, "launches": int(grad_n), "avg_launch_ms": grad_ms / max(grad_n, 1),
                           "avg_rows_per_launch": rows / max(grad_n, 1), "peak_source": peak_src,
                           "kernel_share_of_step": grad_ms / ms,
                           # what the tensor pipe executes for those algorithmic flops: K and N padded to 16, the position
                           # operand in `terms` bf16 terms and the residual in `rterms`
                           "executed_over_algorithmic": exe,
                           "executed": (ach * exe) if ach else None}
    else:
        sz = 8 if (a.config == "c4" and a.dtype == "f64") else 4
        peak_bw = peaks.get("hbm_gbs")
        peak_src = "measured copy bandwidth (MEASURED_PEAKS.json)" if peak_bw else "fallback 6650 (B200_PROFILING.md)"
        peak_bw = peak_bw or 6650.0
        out["dtype"] = "f64" if sz == 8 else ("f32 (bf16 operand splits on tcgen05 for the gradient)" if a.config == "c2" else "f32")
        ach = leap / world / (ms * 1e-3) * 13 * D * sz / 1e9      # per GPU
        # per chain: 14 phase-point slots x (q, p, grad) + 2 x 10 stack vectors + 6 vectors of the main tree / metric
        config["l2"] = "inputs larger than L2 (chain state %d MB per GPU)" % (C * (14 * 3 + 20 + 6) * ((D + 31) // 32 * 32) * sz // 2**20)
        out["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak_bw, "unit": "GB/s", "frac": ach / peak_bw,
                           "traffic": None, "kernel": "k_advance", "peak_source": peak_src,
                           "model": "13*D*sizeof(T) bytes per chain-leapfrog (SURVEY.md section 8d streaming model), per GPU",
                           "gradient_kernel_share_of_step": (grad_ms / ms) if grad_n else 0.0}
        if a.config == "c2" and grad_n:
            out["roofline"]["gradient_alg_tflops"] = 2.0 * D * D * rows / (grad_ms * 1e-3) / 1e12
    if replicas_equal is not None:
        out["replicas_bit_identical"] = replicas_equal
    if min_ess is not None:
        out["min_ess"] = {"value": min_ess, "per_s": min_ess / dt, "unit": "min over coordinates of multi-chain bulk ESS (/s: e2e wall clock)",
                          "draws_per_chain": int(kept.shape[1])}
    if fp64 is not None:
        out["fp64_value"] = fp64
    if world == 1 and not a.no_cpu_baseline:
        v, cores, busy, sample, extra = reference_leg(a, budget_s=15.0)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "busy_threads_measured": round(busy, 1), "kind": "port", "sample": sample}
        out["cpu_baseline"].update(extra)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — leapfrog steps/s of the NUTS hot path on BASELINE config 3
(Bayesian logistic regression N=1e6, D=100, 4096 chains sharded over the GPUs).

  python bench.py --gpus N --steps K --warmup W            our CUDA engine
  python bench.py --impl reference ...                      CPU restatement of the reference on host cores

A "step" is `--transitions` NUTS transitions of every chain (one bnuts_sample call):
momentum refresh, tree building (leapfrog + gradient per leaf), selection, statistics.
Chains run their transitions asynchronously inside the call (a chain starts its next
tree as soon as it finishes one); idle chains occupy no rows of the gradient kernel.
`value` = leapfrog steps (Σ TreeStatisticsNUTS.steps, src/NUTS.jl:238-239) of all
chains on all GPUs ÷ device time of the K timed steps, state resident in HBM.
`e2e` = the same through the C ABI with host buffers: positions are uploaded and
draws + statistics downloaded inside the timed region.
"""
import argparse
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

METRIC = "leapfrog steps/sec (4096 chains, logistic reg N=1e6 D=100)"
UNIT = "leapfrog steps/s"


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


def cpu_leapfrog_rate(bits, y, D, budget_s, seed=1):
    """The oracle (CPU restatement of src/kinetic_energy.jl:126-163 + the logistic target), one chain per
    host thread like Threads.@threads in src/mcmc.jl:150-157, on a bounded number of bare leapfrog steps."""
    import inplacedhmc_jl_b200 as bn
    from conftest import build_oracle
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    lib = bn.load_library(build_oracle())
    e = bn.Engine(cores, D, dtype=bn.F64, lib=lib, seed=seed)
    e.model_logistic(bits, y, 1.0, row_blocks=1)
    rng = np.random.default_rng(seed)
    e.set_positions(rng.normal(size=(cores, D)) * 0.01)
    p = rng.normal(size=(cores, D))
    t = time.perf_counter(); e.leapfrog(p, 1e-3, 1); t1 = time.perf_counter() - t
    n = int(max(1, min(64, budget_s / max(t1, 1e-3))))
    t = time.perf_counter(); e.leapfrog(p, 1e-3, n); dt = time.perf_counter() - t
    return cores * n / dt, cores, f"{n} bare leapfrog steps x {cores} chains (one per host thread), N={len(y)} D={D}, fp64 oracle"


def run_reference(a, rank, world):
    if rank != 0:
        return
    bits, y, _ = synth(a.rows, a.dim)
    rates = []
    sample = ""
    t0 = time.perf_counter()
    for i in range(a.warmup + a.steps):
        r, cores, sample = cpu_leapfrog_rate(bits, y, a.dim, budget_s=max(2.0, 60.0 / (a.warmup + a.steps)))
        if i >= a.warmup:
            rates.append(r)
    val = float(np.mean(rates))
    ms = (time.perf_counter() - t0) * 1e3 / (a.warmup + a.steps)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c3: Bayesian logistic regression N=%d D=%d, 4096 chains total, max_depth 10" % (a.rows, a.dim),
                   "sample": "bounded sample of the workload per step: bare leapfrogs (integrator + gradient) of one chain per host thread"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=4096)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=100)
    ap.add_argument("--adapt", type=int, default=40, help="length of the first/last step-size-only warmup stage (untimed)")
    ap.add_argument("--transitions", type=int, default=64, help="NUTS transitions per chain per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference", action="store_true", help="keep the exact three-term position operand")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (BASELINE config 3: --chains in total, sharded over the GPUs) or weak (--chains per GPU)")
    a = ap.parse_args()
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

    N, D = a.rows, a.dim
    C = a.chains // world if a.scaling == "strong" else a.chains   # chains shard across GPUs, no communication
    bits, y, beta = synth(N, D)
    e = bn.Engine(C, D, dtype=bn.F32, seed=20261018, chain_offset=rank * C, device=local, gradient_path=bn.GRAD_TENSOR)
    e.model_logistic(bits, y, 1.0)
    # ≙ default_warmup_stages (src/warmup.jl:361-372): q0 ~ U[-2,2]^D (:73), FindLocalOptimum, InitialStepsizeSearch, ...
    t_w = time.perf_counter()
    e.set_positions(None)
    e.find_local_optimum(1e-4, 50)
    # reference point of the tensor path (include/bnuts.h) = the optimum just found (across-chain mean); the engine
    # checks it and keeps the exact three-term path if it is refused
    terms = 3
    if not a.no_reference:
        try:
            e.logistic_set_reference(e.get_state()[0].mean(axis=0))
            terms = 2
        except bn.BnutsError as ex:
            print("reference point refused: %s" % ex, file=sys.stderr)
    e.find_initial_stepsize()
    # ≙ default_warmup_stages (src/warmup.jl:361-372) with shorter windows: step size only, then
    # step size + per-chain diagonal metric in doubling windows, then step size only
    if a.adapt > 0:
        for n, mk in ((a.adapt, bn.METRIC_NONE), (25, bn.METRIC_DIAG), (50, bn.METRIC_DIAG), (100, bn.METRIC_DIAG),
                      (a.adapt, bn.METRIC_NONE)):
            e.warmup_stage(n, mk, keep=False)
    t_w = time.perf_counter() - t_w
    T = a.transitions
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
    leap = allsum(c1["leapfrogs"] - c0["leapfrogs"])
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    value = leap / (ms * 1e-3)

    # ---------------- timed region 2: through the C ABI with host buffers
    qh = torch.empty((C, D), dtype=torch.float64).pin_memory().numpy()
    qh[:] = e.get_state()[0]
    chain = torch.empty((C, T, D), dtype=torch.float64).pin_memory().numpy()
    stats = np.zeros((C, T), dtype=bn.TREE_STATS_DTYPE)
    kept = np.empty((C, a.steps * T, D))          # the e2e draws form one continuous chain per chain id
    barrier()
    t0 = time.perf_counter(); leap_e2e = 0
    for k in range(a.steps):
        e.set_positions(qh)                       # H2D of this step's input positions (+ their gradient)
        e.sample(T, out=(chain, stats))           # D2H of the draws and tree statistics
        leap_e2e += int(stats["steps"].sum())
        qh[:] = chain[:, T - 1]
        kept[:, k * T:(k + 1) * T] = chain
    barrier()
    dt = allmax(time.perf_counter() - t0)
    e2e = allsum(leap_e2e) / dt
    # min-ESS/s (second half of BASELINE.json's metric): multi-chain bulk ESS per coordinate of the e2e draws;
    # chains on different GPUs are independent, so per-coordinate ESS adds across ranks
    # (estimated on the first 512 chains of the rank and scaled to all of them: chains are i.i.d. replicas)
    nsub = min(C, 512)
    ess_d = (np.array([bn.diagnostics.ess(kept[:nsub, :, d]) for d in range(D)]) * (C / nsub)) if kept.shape[1] >= 4 else np.zeros(D)
    if dist is not None:
        tt = torch.tensor(ess_d, dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        ess_d = tt.cpu().numpy()
    min_ess = float(ess_d.min())

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
    peak_tf = peaks.get("bf16_tflops_sustained")
    peak_src = "measured sustained (MEASURED_PEAKS.json)" if peak_tf else "fallback 1400 (B200_PROFILING.md)"
    peak_tf = peak_tf or 1400.0
    rows = c1["gradient_rows"] - c0["gradient_rows"]   # active chains summed over the gradient launches
    # algorithmic flops: X·B and Xᵀ·R with the true D, 2 flop per MAC, only rows that were requested
    ach = 4.0 * N * D * rows / (grad_ms * 1e-3) / 1e12 if grad_n else None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": "f32 (exact bf16 operand splits on tcgen05, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": "c3: Bayesian logistic regression N=%d D=%d, %d chains total, max_depth 10" % (N, D, C * world),
                   "chains_per_gpu": C, "parallelism": "chains sharded, no collective",
                   "l2": "inputs larger than L2 (X is %d MB bf16)" % (N * 128 * 2 // 2**20),
                   "init": "q0 ~ U[-2,2]^D; untimed warmup: FindLocalOptimum(1e-4, 50), step size search, stages %d|25,50,100 (diag metric)|%d, %.1f s" % (a.adapt, a.adapt, t_w),
                   "step": "%d NUTS transitions of every chain (async within the call)" % T,
                   "position_operand_terms": terms,
                   "mean_tree_depth": float(stats["depth"].mean()), "mean_leapfrogs_per_transition": float(stats["steps"].mean()),
                   "lockstep_steps_timed": int(c1["lockstep_steps"] - c0["lockstep_steps"]),
                   "active_row_fraction": float(leap / max(1, (c1["lockstep_steps"] - c0["lockstep_steps"]) * C * world))},
        "gpu_launches": int(launches),
        "leapfrogs_timed": int(leap),
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (ach / peak_tf) if ach else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture
                     # (profiles/r1_ncu_full_details_k_logistic_tc_4096rows.csv: 292.4 MB + 14.5 MB; X is read once per
                     # launch whatever the number of active rows: 256 MB algorithmic)
                     "traffic": 306.9e6 if (N == 1_000_000 and D == 100) else None, "traffic_unit": "bytes per launch (ncu, full 4096-row launch)",
                     "kernel": "k_logistic_tc",
                     "launches": int(grad_n), "avg_launch_ms": grad_ms / max(grad_n, 1), "avg_rows_per_launch": rows / max(grad_n, 1), "peak_source": peak_src,
                     "kernel_share_of_step": grad_ms / ms,
                     # what the tensor pipe executes for those algorithmic flops: K and N padded to 16 (D + 3 reference
                     # columns -> dk), the position operand in `terms` bf16 terms and the residual in two
                     "executed_over_algorithmic": (terms + 2) * (-(-(D + 3) // 16) * 16) / (2.0 * D),
                     "executed": (ach * (terms + 2) * (-(-(D + 3) // 16) * 16) / (2.0 * D)) if ach else None},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(qh.nbytes), "d2h_bytes_per_step": int(chain.nbytes + stats.nbytes)},
        "clocks": clk,
        "min_ess": {"value": min_ess, "per_s": min_ess / dt, "unit": "min over coordinates of multi-chain bulk ESS (/s: e2e wall clock)",
                    "draws_per_chain": int(kept.shape[1])},
    }
    if world == 1 and not a.no_cpu_baseline:
        v, cores, sample = cpu_leapfrog_rate(bits, y, D, budget_s=15.0)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

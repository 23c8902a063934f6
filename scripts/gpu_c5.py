"""Row-sharded logistic regression (BASELINE config 5 shape at a reduced N; D up to 256): rows of X split over the
ranks, all chains replicated, one NCCL all-reduce of the folded gradient partials per leapfrog step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/gpu_c5.py --rows 2000000

Checks (rank 0 prints one JSON line): replicas bit-identical across ranks; gradient equal to a single engine
holding all rows (tolerance); lockstep step time and the share of the all-reduce."""
import argparse, hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import inplacedhmc_jl_b200 as bn
from bench import synth

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2_000_000)
ap.add_argument("--dim", type=int, default=100)
ap.add_argument("--chains", type=int, default=4096)
ap.add_argument("--transitions", type=int, default=8)
ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"])
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, D, C = a.rows, a.dim, a.chains
bits, y, beta = synth(N, D)
lo, hi = rank * N // world, (rank + 1) * N // world
e = bn.Engine(C, D, dtype=bn.F32, seed=20261018, device=local, gradient_path=bn.GRAD_TENSOR)
e.model_logistic(bits[lo:hi], y[lo:hi], 1.0)
if a.exchange == "nccl":
    ids = [bn.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    e.set_nccl(ids[0], world, rank)
else:   # the library's own kernels over peer memory (NVLink stores + flags)
    hs = [None] * world
    dist.all_gather_object(hs, e.p2p_export())
    e.p2p_connect(hs, rank)
rng = np.random.default_rng(7)
q0 = np.asarray(beta[None, :] + rng.normal(size=(C, D)) * 2e-3, dtype=np.float32).astype(np.float64)
# ≙ FindLocalOptimum: a few ascent steps from the data-generating point give the mode = the reference point of the
# tensor path (for D > 128 the kernel always works around a reference)
e.set_positions(q0)
e.find_local_optimum(1e-4, 25)
bref = e.get_state()[0].mean(axis=0)
e.logistic_set_reference(bref)
e.set_positions(q0)
_, g, l = e.get_state()
out = {"exchange": a.exchange, "world": world, "rows_total": N, "rows_per_rank": hi - lo, "dim": D, "chains": C}
if rank == 0:   # the same chains on one engine that holds every row
    f = bn.Engine(C, D, dtype=bn.F32, seed=20261018, device=local, gradient_path=bn.GRAD_TENSOR)
    f.model_logistic(bits, y, 1.0); f.logistic_set_reference(bref); f.set_positions(q0)
    _, gf, lf = f.get_state()
    out["grad_rel_vs_single_engine"] = float(np.max(np.linalg.norm(g - gf, axis=1) / np.linalg.norm(gf, axis=1)))
    out["logdensity_rel_vs_single_engine"] = float(np.max(np.abs(l - lf) / np.abs(lf)))
    f.close()
e.find_initial_stepsize()
e.warmup_stage(10, bn.METRIC_NONE, keep=False)
c0 = e.counters(); torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
ch, st = e.sample(a.transitions)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
c1 = e.counters()
h = hashlib.sha256(ch.tobytes() + st.tobytes()).hexdigest()
hs = [None] * world
dist.all_gather_object(hs, h)
tmax = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
if rank == 0:
    steps = c1["lockstep_steps"] - c0["lockstep_steps"]; leap = c1["leapfrogs"] - c0["leapfrogs"]
    out.update({"replicas_bit_identical": len(set(hs)) == 1, "lockstep_steps": steps, "leapfrogs": leap,
                "seconds": float(tmax[0]), "ms_per_lockstep_step": 1e3 * float(tmax[0]) / max(steps, 1),
                "chain_leapfrogs_per_s": leap / float(tmax[0]),
                "alg_TFLOPs_total": 4.0 * N * D * (c1["gradient_rows"] - c0["gradient_rows"]) / float(tmax[0]) / 1e12,
                "exchange_bytes_per_step_full": C * (((D + 31) // 32 * 32) * 4 + 8)})
    print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()

"""c5 set-up path of bench.py on one GPU at a reduced row count: state of the chains after every phase."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import inplacedhmc_jl_b200 as bn
N = int(os.environ.get("NROWS", 2_000_000)); D = 256; C = int(os.environ.get("CHAINS", 4096))
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:   # rows sharded: torchrun --nproc-per-node 2 scripts/gpu_c5_debug.py
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
e = bn.Engine(C, D, dtype=bn.F32, seed=20261018, device=local, gradient_path=bn.GRAD_TENSOR)
e.model_logistic_synthetic(5, N * rank // world, N * (rank + 1) // world - N * rank // world, 1.0)
if world > 1:
    ids = [bn.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    e.set_nccl(ids[0], world, rank)
_, _, beta = bn.synth_logistic_rows(5, 0, 0, D)
def show(tag):
    q, g, l = e.get_state()
    st = e.chain_status()
    if rank != 0:
        return
    print(tag, "status counts", dict(zip(*np.unique(st, return_counts=True))), "|q - beta*| median %.3g max %.3g" % (np.median(np.linalg.norm(q - beta, axis=1)), np.linalg.norm(q - beta, axis=1).max()),
          "|grad| median %.3g max %.3g" % (np.median(np.linalg.norm(g, axis=1)), np.nanmax(np.linalg.norm(g, axis=1))), "l min %.6g max %.6g nonfinite %d" % (np.nanmin(l), np.nanmax(l), (~np.isfinite(l)).sum()),
          "spread of chains about their mean: %.3g" % np.linalg.norm(q - q.mean(0), axis=1).max(), flush=True)
e.set_positions(None); show("init")
for r in range(3):
    try:
        e.find_local_optimum(1e-4, int(os.environ.get("OPT_ITERS", 30)))
    except bn.BnutsError as ex:
        print("optimum:", ex)
    show("after optimum %d" % r)
    e.logistic_set_reference(e.get_state()[0].mean(axis=0)); show("after set_reference %d" % r)
try:
    e.find_initial_stepsize()
except bn.BnutsError as ex:
    print("search:", ex)
show("after search")
eps = e.get_stepsize()
if os.environ.get("CHECK") == "1":   # the reference values of a single engine with all rows (rank 0 only builds it)
    q = e.get_state()[0]
    qs = q[:64] + np.random.default_rng(3).normal(size=(64, D)) * 1e-3
    qq = np.tile(qs, (C // 64, 1))
    e.set_positions(qq); _, g, l = e.get_state()
    if rank == 0:
        f = bn.Engine(C, D, dtype=bn.F32, seed=20261018, device=local, gradient_path=bn.GRAD_TENSOR)
        f.model_logistic_synthetic(5, 0, N, 1.0); f.logistic_set_reference(q.mean(axis=0)); f.set_positions(qq)
        _, gf, lf = f.get_state()
        print("rows sharded vs one engine: grad rel diff max %.3g, log density abs diff max %.3g (l = %.8g)" % (
            np.max(np.linalg.norm(g - gf, axis=1) / np.linalg.norm(gf, axis=1)), np.max(np.abs(l - lf)), lf[0]), flush=True)
if rank == 0:
    print("eps", np.percentile(eps, [0, 50, 100]))
if world > 1:
    dist.barrier(); dist.destroy_process_group()

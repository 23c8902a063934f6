"""world_size-2 (gloo, CPU) check of the multi-GPU plan of SURVEY.md §8(e): chains shard by global id with no
data-path collective; ranks only synchronise and reduce the timing / leapfrog totals, as bench.py does."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, HOSTEMU_SO, build_hostemu


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import inplacedhmc_jl_b200 as bn
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, D = 6, 8
    e = bn.Engine(C // world, D, max_depth=5, lib=HOSTEMU_SO, seed=3, chain_offset=rank * (C // world))
    e.model_funnel(); e.set_positions(None); e.set_stepsize(0.3)
    dist.barrier()
    ch, st = e.sample(20)
    leap = torch.tensor([float(st["steps"].sum())], dtype=torch.float64)
    dist.all_reduce(leap, op=dist.ReduceOp.SUM)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out, f"chain{rank}.npy"), ch)
    np.save(os.path.join(out, f"tot{rank}.npy"), np.array([leap.item(), t.item()]))
    dist.destroy_process_group()


def test_two_ranks_equal_one_engine(tmp_path, bn):
    build_hostemu()
    world = 2
    mp.spawn(_worker, args=(world, 29517, str(tmp_path)), nprocs=world, join=True)
    e = bn.Engine(6, 8, max_depth=5, lib=HOSTEMU_SO, seed=3)
    e.model_funnel(); e.set_positions(None); e.set_stepsize(0.3)
    ch, st = e.sample(20)
    got = np.concatenate([np.load(tmp_path / f"chain{r}.npy") for r in range(world)])
    assert got.tobytes() == ch.tobytes()
    tot = np.load(tmp_path / "tot0.npy")
    assert tot[0] == st["steps"].sum() and tot[1] == 2.0


# ---------------------------------------------------------------- row-sharded data
def _gloo_allreduce(addr, count, dtype):
    import ctypes
    ct = ctypes.c_double if dtype == 0 else ctypes.c_float
    a = np.ctypeslib.as_array((ct * count).from_address(addr))
    t = torch.from_numpy(a)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)       # in place: t shares memory with the engine's buffer


def _row_worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import inplacedhmc_jl_b200 as bn
    from conftest import make_logistic
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, D, C = 400, 9, 7
    X, y, beta = make_logistic(N, D, seed=21)
    lo, hi = rank * N // world, (rank + 1) * N // world
    e = bn.Engine(C, D, max_depth=6, lib=HOSTEMU_SO, seed=5)          # all chains on every rank, same seed
    e.model_logistic(X[lo:hi], y[lo:hi], 1.0, row_blocks=2)            # one shard of the rows
    e.set_allreduce(_gloo_allreduce)
    rng = np.random.default_rng(4)
    e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 0.3)
    q, g, l = e.get_state()
    e.find_initial_stepsize()
    ch0, st0, _ = e.warmup_stage(30, 1)
    ch, st, sel = e.sample(25, want_index=True)
    np.savez(os.path.join(out, f"rows{rank}.npz"), g=g, l=l, ch=ch, st=st, sel=sel, eps=e.get_stepsize(), lockstep=e.counters()["lockstep_steps"])
    dist.destroy_process_group()


def test_row_sharded_replicas_agree_and_match_one_engine(tmp_path, bn):
    """Two ranks hold half the rows each and all chains; gradients are summed per leapfrog (gloo).  The replicas
    must be bit-identical to each other (same tree decisions on every rank) and equal to a single engine that
    holds all rows up to summation order."""
    from conftest import make_logistic
    build_hostemu()
    world = 2
    mp.spawn(_row_worker, args=(world, 29519, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rows0.npz"), np.load(tmp_path / "rows1.npz")
    for k in ("g", "l", "ch", "st", "sel", "eps"):
        assert r0[k].tobytes() == r1[k].tobytes(), k                   # replica determinism, bit for bit
    N, D, C = 400, 9, 7
    X, y, beta = make_logistic(N, D, seed=21)
    e = bn.Engine(C, D, max_depth=6, lib=HOSTEMU_SO, seed=5)
    e.model_logistic(X, y, 1.0, row_blocks=2)
    rng = np.random.default_rng(4)
    e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 0.3)
    q, g, l = e.get_state()
    np.testing.assert_allclose(r0["g"], g, rtol=1e-12, atol=1e-12)     # same sum, different association
    np.testing.assert_allclose(r0["l"], l, rtol=1e-13)
    e.find_initial_stepsize()
    e.warmup_stage(30, 1)
    ch, st, sel = e.sample(25, want_index=True)
    # free-running chains: rounding differences of 1e-16 can flip a decision eventually; the bulk must coincide
    same = (st["depth"] == r0["st"]["depth"]) & (st["steps"] == r0["st"]["steps"]) & (sel == r0["sel"])
    assert same[:, :3].all() and same.mean() > 0.8, same.mean()
    ok = same.all(axis=1)
    np.testing.assert_allclose(r0["ch"][ok], ch[ok], rtol=1e-8, atol=1e-8)


# ---------------------------------------------------------------- config c5's data flow: every rank GENERATES its shard
def _synth_row_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import inplacedhmc_jl_b200 as bn
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, D, C, seed = 600, 10, 5, 5
    lo, hi = rank * N // world, (rank + 1) * N // world
    e = bn.Engine(C, D, max_depth=6, lib=HOSTEMU_SO, seed=9)
    e.model_logistic_synthetic(seed, lo, hi - lo, 1.0, row_blocks=2)   # rows [lo, hi) of the one conceptual matrix
    e.set_allreduce(_gloo_allreduce)
    _, _, beta = bn.synth_logistic_rows(seed, 0, 0, D, lib=bn.load_library(HOSTEMU_SO))
    rng = np.random.default_rng(6)
    e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 0.3)
    q, g, l = e.get_state()
    e.set_stepsize(0.05)
    ch, st = e.sample(10)
    np.savez(os.path.join(out, f"synth{rank}.npz"), g=g, l=l, ch=ch, st=st)
    dist.destroy_process_group()


def test_row_sharded_synthetic_rows(tmp_path, bn):
    """Rows generated where they are used (bnuts_model_logistic_synthetic with a row offset per rank) add up to the
    model of a single engine that generates all rows: the sharding does not change the matrix."""
    build_hostemu()
    world = 2
    mp.spawn(_synth_row_worker, args=(world, 29523, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "synth0.npz"), np.load(tmp_path / "synth1.npz")
    for k in ("g", "l", "ch", "st"):
        assert r0[k].tobytes() == r1[k].tobytes(), k
    N, D, C, seed = 600, 10, 5, 5
    e = bn.Engine(C, D, max_depth=6, lib=HOSTEMU_SO, seed=9)
    e.model_logistic_synthetic(seed, 0, N, 1.0, row_blocks=2)
    _, _, beta = bn.synth_logistic_rows(seed, 0, 0, D, lib=bn.load_library(HOSTEMU_SO))
    rng = np.random.default_rng(6)
    e.set_positions(beta[None, :] + rng.normal(size=(C, D)) * 0.3)
    q, g, l = e.get_state()
    np.testing.assert_allclose(r0["g"], g, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(r0["l"], l, rtol=1e-13)

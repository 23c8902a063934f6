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

"""The N>1 path on CPU: world_size-2 gloo.  LUNs shard one per rank with no data-path collective
(SURVEY.md 8(e)); the only exchanges are the barrier and the max-over-ranks of the timed region."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    import util
    from oim_b200 import traces
    from oracle import bindings
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    plan = bench.shard_plan(rank, world)
    # every rank owns ONE LUN: its own store, its own trace (seeded by rank), nothing shared
    nb = 32768
    t = traces.partitioned_queues(4, 64, nb, pattern="randrw", read_pct=70, seed=plan["trace_seed"])
    dist.barrier()
    cpls, arena, store = util.run_oracle(bindings.PortOracle, t, nb, store_seed=plan["store_seed"])
    # stand-in for the device-timed region: rank r "takes" (r+1) ms; the job time is the max
    my_ms = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(my_ms, op=dist.ReduceOp.MAX)
    units = torch.tensor([len(t)], dtype=torch.int64)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    names = [None] * world
    dist.all_gather_object(names, (plan["bdev"], plan["ctrlr"], util.sha(store)))
    value = bench.aggregate(len(t), steps=1, world=world, max_ms=float(my_ms))
    if rank == 0:
        np.save(os.path.join(out_dir, "r.npy"), np.array([float(my_ms), int(units), value]))
        with open(os.path.join(out_dir, "names.txt"), "w") as f:
            f.write(repr(names))
    assert (cpls["status"] == 0).all()
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding(oracles, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    max_ms, units, value = np.load(tmp_path / "r.npy")
    assert max_ms == 2.0                      # max over ranks, not rank 0's own time
    assert units == 2 * 256                   # weak scaling: per-rank work is fixed
    assert value == pytest.approx(512 / 2e-3)  # whole-job units / slowest rank
    names = eval((tmp_path / "names.txt").read_text())
    assert len({n[0] for n in names}) == world and len({n[1] for n in names}) == world   # distinct LUNs
    assert names[0][2] != names[1][2]         # different traces/seeds -> different stores: nothing shared


def test_shard_plan_is_one_lun_per_rank():
    sys.path.insert(0, ROOT)
    import bench
    plans = [bench.shard_plan(r, 8) for r in range(8)]
    assert len({p["bdev"] for p in plans}) == 8 and len({p["trace_seed"] for p in plans}) == 8
    assert all(p["target"] == 0 for p in plans)
    with pytest.raises(AssertionError):
        bench.shard_plan(8, 8)

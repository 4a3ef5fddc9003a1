"""How much can the box's host side take when N GPUs write into host memory at once?  Run under torchrun, one rank
per GPU (gpurun --gpus N -- python -m torch.distributed.run --nproc-per-node N tools/pcie_concurrent_probe.py):
every rank measures D2H by the copy engine and by SM stores, all ranks concurrently, first as the scheduler placed
it, then bound to the GPU's NUMA node.  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oim_b200 import build, hostmem, lib  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
torch.zeros(1, device="cuda")
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
build.build()
lib.init([local])


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gather(p):
    t = torch.tensor([p["copy_engine_gbs"], p["sm_stores_gbs"]], dtype=torch.float64, device="cuda")
    if world == 1:
        return [[float(t[0]), float(t[1])]]
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [[float(x[0]), float(x[1])] for x in out]


res = {"gpus": world}
for phase in ("unbound", "bound"):
    info = hostmem.bind_to_gpu_node(local) if phase == "bound" else {"node": hostmem.gpu_numa_node(local)}
    per = gather(hostmem.concurrent_d2h_probe(torch, local, barrier, seconds=1.0))
    res[phase] = {"copy_engine_total_gbs": sum(p[0] for p in per), "sm_stores_total_gbs": sum(p[1] for p in per),
                  "copy_engine_per_gpu": [round(p[0], 1) for p in per], "sm_stores_per_gpu": [round(p[1], 1) for p in per],
                  "rank0_placement": info}
if rank == 0:
    print(json.dumps(res))
lib.fini()
if world > 1:
    dist.destroy_process_group()

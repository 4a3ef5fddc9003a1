"""BASELINE config 5 on N GPUs (gpurun --gpus N; N = 2 or 4): N two-way mirrored 8 GiB bdevs, bdev i with its
primary replica on GPU i and the second replica on GPU (i+1) % N, all written at once with 128 KiB sequential
WRITEs (32 x 4 KiB SG pages).  Every GPU is primary of one mirror and replica holder of another, so each
NVLink port carries one payload out and one payload in.  One process, one session per bdev on its own stream;
timing by CUDA events on each session's stream, max over the sessions.  Prints one JSON line."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oim_b200 import abi, build, lib, traces  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
NB, nq, per_q, steps, warmup = 16777216, 256, 256, 10, 3
build.build()
for d in range(n):
    torch.zeros(1, device=f"cuda:{d}")
lib.init(list(range(n)))
t = traces.uniform_trace(nq * per_q, NB, io_blocks=256, pattern="seqwrite", sg="pages", seed=9)
sessions = []
for i in range(n):
    torch.cuda.set_device(i)
    arena = torch.empty(t.arena_bytes, dtype=torch.uint8, device=f"cuda:{i}")
    arena.view(torch.int64)[:] = 0x1122334455667788 + i
    d_reqs = torch.from_numpy(t.reqs.view(np.uint8)).to(f"cuda:{i}")
    d_iovs = torch.from_numpy(t.bind(arena.data_ptr()).view(np.uint8)).to(f"cuda:{i}")
    d_cpls = torch.zeros(len(t.reqs) * 48, dtype=torch.uint8, device=f"cuda:{i}")
    name = lib.construct_mirror_bdev(NB, 512, [i, (i + 1) % n], name=f"mirror{i}")
    lib.construct_vhost_scsi_controller(f"m{i}.ctl")
    lib.add_vhost_scsi_lun(f"m{i}.ctl", 0, name)
    lun = lib.Lun(f"m{i}.ctl", 0, num_queues=nq, queue_size=32)
    sessions.append((lun, lib.Timer(), arena, d_reqs, d_iovs, d_cpls, name))


def step():
    for lun, _, _, d_reqs, d_iovs, d_cpls, _ in sessions:
        lun.submit_batch(nq, per_q, d_reqs.data_ptr(), d_iovs.data_ptr(), len(t.iovs), d_cpls.data_ptr(), abi.MEM_DEVICE)


for _ in range(warmup):
    step()
for s in sessions:
    s[0].sync()
for lun, timer, *_ in sessions:
    timer.start(lun)
for _ in range(steps):
    step()
for lun, timer, *_ in sessions:
    timer.stop(lun)
for s in sessions:
    s[0].sync()
ms = [s[1].elapsed_ms() / steps for s in sessions]
for lun, _, _, _, _, d_cpls, name in sessions:
    c = np.frombuffer(d_cpls.cpu().numpy().tobytes(), dtype=abi.cpl_dtype)
    assert not c["status"].any()
    a = lib.bdev_read_raw(name, 0, 1 << 20, replica=0)
    b = lib.bdev_read_raw(name, 0, 1 << 20, replica=1)
    assert (a == b).all() and a.any(), "replicas differ"
per = [len(t.reqs) * 131072 / m / 1e6 for m in ms]
print(json.dumps({"gpus": n, "mirrors": n, "layout": "bdev i: replica 0 on GPU i, replica 1 on GPU (i+1)%N",
                  "ms_per_pass_max": max(ms), "user_gbs_per_mirror": per,
                  "user_gbs_total": n * len(t.reqs) * 131072 / max(ms) / 1e6,
                  "nvlink_gbs_per_gpu_each_direction": len(t.reqs) * 131072 / max(ms) / 1e6,
                  "note": "user GB/s = payload written once by the guest; every byte also lands in the peer's HBM over NVLink"}))
for s in sessions:
    s[0].close()

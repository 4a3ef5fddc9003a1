"""4 KiB random read, everything resident in HBM, as a function of the number of request queues (and through
guest virtio rings): what a guest with 1..254 queues gets.  Run under gpurun; prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oim_b200 import abi, build, lib, traces, vring  # noqa: E402

NB, BLOCK = 16777216, 512
build.build()
torch.zeros(1, device="cuda:0")
lib.init([0])
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
name = lib.construct_malloc_bdev(NB, BLOCK, name="sweep0", device=0)
lib.construct_vhost_scsi_controller("sweep.ctl")
lib.add_vhost_scsi_lun("sweep.ctl", 0, name)
timer = lib.Timer()
out = {"slots": {}, "vring": {}, "mixed": {}, "write": {}}
total = int(os.environ.get("SWEEP_TOTAL", 1 << 20))
steps = 5
qs = [int(x) for x in os.environ.get("SWEEP_QUEUES", "1,2,4,8,16,64,254,1024").split(",")]
with lib.Lun("sweep.ctl", 0, num_queues=1024, queue_size=32) as lun:
    for pattern, key in (("randread", "slots"), ("randrw", "mixed"), ("randwrite", "write")):
        for nq in qs:
            per_q = max(32, total // nq // 32 * 32)
            n = nq * per_q
            if pattern == "randrw":
                t = traces.partitioned_queues(nq, per_q, NB, pattern="randrw", read_pct=70, io_blocks=8, seed=5)
            else:
                t = traces.uniform_trace(n, NB, io_blocks=8, pattern=pattern, seed=3)
            arena = torch.zeros(t.arena_bytes, dtype=torch.uint8, device="cuda")
            d_reqs = torch.from_numpy(t.reqs.view(np.uint8)).cuda()
            d_iovs = torch.from_numpy(t.bind(arena.data_ptr()).view(np.uint8)).cuda()
            d_cpls = torch.zeros(n * 48, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            for _ in range(2):
                lun.submit_batch(nq, per_q, d_reqs.data_ptr(), d_iovs.data_ptr(), len(t.iovs), d_cpls.data_ptr(), abi.MEM_DEVICE)
            lun.sync()
            timer.start(lun)
            for _ in range(steps):
                lun.submit_batch(nq, per_q, d_reqs.data_ptr(), d_iovs.data_ptr(), len(t.iovs), d_cpls.data_ptr(), abi.MEM_DEVICE)
            timer.stop(lun)
            lun.sync()
            ms = timer.elapsed_ms() / steps
            c = np.frombuffer(d_cpls.cpu().numpy().tobytes(), dtype=abi.cpl_dtype)
            assert not c["status"].any() and (c["used_len"] > 0).all()
            iops = n / ms * 1e3
            out[key][nq] = {"iops": iops, "hbm_frac": 2 * 4096 * iops / 1e9 / peak, "ms": ms, "requests": n}
            del arena, d_reqs, d_iovs, d_cpls
            torch.cuda.empty_cache()
        print(json.dumps({key: out[key]}), file=sys.stderr, flush=True)
# guest virtio rings in HBM
for nq in ([] if os.environ.get("SWEEP_NO_VRING") else [q for q in qs if q <= 254] + [4096]):
    per, ring = 256, 1024            # 3-descriptor chains: a 1024-entry ring (SPDK_VHOST_MAX_VQ_SIZE) holds 341 requests
    g = vring.build_uniform_queues(nq, per, NB, ring_size=ring, seed=11)
    guest = torch.empty(g.total_bytes(), dtype=torch.uint8, device="cuda")
    guest[:g.data_off] = torch.from_numpy(g.arena).cuda()
    guest[g.data_off:].zero_()
    gbase = guest.data_ptr()
    with lib.Lun("sweep.ctl", 0, num_queues=nq, queue_size=32) as vlun:
        vlun.set_mem_table(np.array([g.gpa_base, g.total_bytes(), gbase], dtype=np.uint64))
        for q in range(nq):
            qb = gbase + q * g.q_stride
            vlun.vq_attach(q, qb + g.desc_off, qb + g.avail_off, qb + g.used_off, ring, 0, 0)
        avail_idx = guest[:g.data_off].view(torch.int16).view(nq, g.q_stride // 2)[:, (g.avail_off + 2) // 2]
        used_idx = guest[:g.data_off].view(torch.int16).view(nq, g.q_stride // 2)[:, (g.used_off + 2) // 2]
        vt, rounds = 0.0, 8
        for k in range(rounds + 2):
            avail_idx.add_(per)
            torch.cuda.current_stream().synchronize()
            timer.start(vlun)
            vlun.vq_kick()
            timer.stop(vlun)
            vlun.sync()
            if k >= 2:
                vt += timer.elapsed_ms()
        want = (per * (rounds + 2)) & 0xFFFF
        u = used_idx.to(torch.int32) & 0xFFFF
        assert int(u.min()) == want and int(u.max()) == want, (int(u.min()), int(u.max()), want)
        iops = nq * per * rounds / vt * 1e3
        out["vring"][nq] = {"iops": iops, "hbm_frac": 2 * 4096 * iops / 1e9 / peak, "ms_per_kick": vt / rounds, "requests_per_kick": nq * per}
    del guest
    torch.cuda.empty_cache()
lib.fini()
print(json.dumps(out))

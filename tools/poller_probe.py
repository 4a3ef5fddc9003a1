"""virtqueue mode with guest memory in pinned host RAM: one launch per kick vs the resident poller (diagnostics)"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from oim_b200 import lib, vring, build
build.build(); lib.load(); lib.init([0])
nq, per_q, ring, nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 256, 1024, 1 << 21
lib.construct_malloc_bdev(nb, 512, name="P0", device=0)
lib.construct_vhost_scsi_controller("pc"); lib.add_vhost_scsi_lun("pc", 0, "P0")
g = vring.build_uniform_queues(nq, per_q, nb, ring_size=ring, seed=3)
guest = torch.empty(g.total_bytes(), dtype=torch.uint8).pin_memory()
guest[:g.data_off] = torch.from_numpy(g.arena)
base = guest.data_ptr()
mem = guest.numpy()
avail = [mem[q * g.q_stride + g.avail_off + 2:q * g.q_stride + g.avail_off + 4].view("<u2") for q in range(nq)]
used = [mem[q * g.q_stride + g.used_off + 2:q * g.q_stride + g.used_off + 4].view("<u2") for q in range(nq)]
for mode in ("kick", "poller", "poller"):
    with lib.Lun("pc", 0, num_queues=nq, queue_size=32) as lun:
        lun.set_mem_table(np.array([g.gpa_base, g.total_bytes(), base], dtype=np.uint64))
        cur = int(used[0][0])
        for q in range(nq):
            qb = base + q * g.q_stride
            lun.vq_attach(q, qb + g.desc_off, qb + g.avail_off, qb + g.used_off, ring, cur, cur)
        if mode == "poller":
            lun.start_poller(idle_timeout_ms=20000)
        ts = []
        for k in range(8):
            want = (cur + per_q * (k + 1)) & 0xFFFF
            t0 = time.perf_counter()
            for a in avail: a[0] = want
            if mode == "kick":
                lun.vq_kick(); lun.sync()
            else:
                while not all(int(u[0]) == want for u in used):
                    if len(sys.argv) > 2: time.sleep(0.0002)      # do not spin on the used indices
                    if time.perf_counter() - t0 > 20: raise SystemExit("timeout")
            ts.append((time.perf_counter() - t0) * 1e3)
        if mode == "poller": lun.stop_poller()
        for q in range(nq): lun.vq_detach(q)
        print(mode, nq, "queues:", " ".join(f"{t:.2f}" for t in ts), "ms", flush=True)

"""Where does the e2e leg lose PCIe bandwidth?  Variants of one 131072 x 4 KiB random-read step with
payload going to pinned host memory: (a) rings in mapped host memory (the e2e path), (b) requests /
SG / completions resident in HBM (only payload crosses PCIe)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oim_b200 import abi, build, lib, traces  # noqa: E402

NB = 16777216
build.build()
torch.zeros(1, device="cuda")
lib.init([0])
lib.construct_malloc_bdev(NB, 512, name="p0", device=0)
lib.construct_vhost_scsi_controller("p.ctl")
lib.add_vhost_scsi_lun("p.ctl", 0, "p0")
nq, per_q = 256, 512
n = nq * per_q
t = traces.uniform_trace(n, NB, io_blocks=8, pattern="randread", seed=5)
host = torch.empty(t.arena_bytes, dtype=torch.uint8).pin_memory()
iovs = t.bind(host.data_ptr())
cpls = np.zeros(n, dtype=abi.cpl_dtype)
timer = lib.Timer()
L = lib.load()
with lib.Lun("p.ctl", 0, num_queues=nq, queue_size=1024) as lun:
    def dev_ms(fn, reps=5):
        best = 1e9
        for _ in range(reps):
            timer.start(lun)
            fn()
            timer.stop(lun)
            lun.sync()
            best = min(best, timer.elapsed_ms())
        return best

    def a():
        L.oimgpu_submit_batch(lun.h, nq, per_q, t.reqs.ctypes.data, iovs.ctypes.data, len(iovs), cpls.ctypes.data, abi.MEM_HOST)
        lun.sync()
        for q in range(nq):
            lun.poll(q, per_q, wait=False)
    ms = dev_ms(a)
    print(f"(a) rings in host memory          : {ms:7.3f} ms  {n * 4096 / ms / 1e6:6.2f} GB/s payload")
    d_reqs = torch.from_numpy(t.reqs.view(np.uint8)).cuda()
    d_iovs = torch.from_numpy(iovs.view(np.uint8)).cuda()
    d_cpls = torch.zeros(n * 48, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def b():
        lun.submit_batch(nq, per_q, d_reqs.data_ptr(), d_iovs.data_ptr(), len(iovs), d_cpls.data_ptr(), abi.MEM_DEVICE)
    ms = dev_ms(b)
    print(f"(b) requests/SG/completions in HBM: {ms:7.3f} ms  {n * 4096 / ms / 1e6:6.2f} GB/s payload")
    for nq2 in (32, 64, 128):
        def c():
            lun.submit_batch(nq2, n // nq2 if n // nq2 <= 1 << 20 else 4096, d_reqs.data_ptr(), d_iovs.data_ptr(), len(iovs), d_cpls.data_ptr(), abi.MEM_DEVICE)
        ms = dev_ms(c)
        print(f"(c) as (b) with {nq2:4d} queues         : {ms:7.3f} ms  {n * 4096 / ms / 1e6:6.2f} GB/s payload")

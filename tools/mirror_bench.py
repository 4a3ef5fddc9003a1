"""Config 5 measurement (run with `gpurun --gpus 2`): 128 KiB sequential write GB/s on
  (a) a plain bdev                       (b) a 2-way mirror, fan-out fused into the write kernel (P2P stores)
  (c) baseline: plain write, then the written extent copied to the peer with cudaMemcpyPeerAsync
      (what a separate collective step - ncclSend/Recv or ncclBroadcast - does at best)
Prints one JSON line."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oim_b200 import abi, build, lib, traces  # noqa: E402

NB = 16777216
build.build()
for d in (0, 1):
    torch.zeros(1, device=f"cuda:{d}")
lib.init([0, 1])
nq, per_q = 256, 256
t = traces.uniform_trace(nq * per_q, NB, io_blocks=256, pattern="seqwrite", sg="pages", seed=9)
torch.cuda.set_device(0)
arena = torch.empty(t.arena_bytes, dtype=torch.uint8, device="cuda:0")
arena.view(torch.int64)[:] = 0x1122334455667788
d_reqs = torch.from_numpy(t.reqs.view(np.uint8)).cuda()
d_iovs = torch.from_numpy(t.bind(arena.data_ptr()).view(np.uint8)).cuda()
d_cpls = torch.zeros(len(t.reqs) * 48, dtype=torch.uint8, device="cuda:0")
out = {}
timer = lib.Timer()
for label, devs in (("plain", [0]), ("mirror2_fused_p2p", [0, 1])):
    name = lib.construct_mirror_bdev(NB, 512, devs, name=label)
    lib.construct_vhost_scsi_controller(label + ".ctl")
    lib.add_vhost_scsi_lun(label + ".ctl", 0, name)
    with lib.Lun(label + ".ctl", 0, num_queues=nq, queue_size=32) as lun:
        def step():
            lun.submit_batch(nq, per_q, d_reqs.data_ptr(), d_iovs.data_ptr(), len(t.iovs), d_cpls.data_ptr(), abi.MEM_DEVICE)
        for _ in range(3):
            step()
        lun.sync()
        timer.start(lun)
        for _ in range(10):
            step()
        timer.stop(lun)
        lun.sync()
        ms = timer.elapsed_ms() / 10
        out[label] = {"ms_per_pass": ms, "gbs": len(t.reqs) * 131072 / ms / 1e6}
        if label == "plain":
            # baseline (c): same write, then ship the 8 GiB extent to the peer as a second step
            peer = torch.empty(NB * 512, dtype=torch.uint8, device="cuda:1")
            src_ptr = lib.get_bdevs(name)[0]["device_ptr"]
            import ctypes
            cudart = ctypes.CDLL("libcudart.so.12")
            torch.cuda.synchronize(0)
            t0 = time.perf_counter()
            for _ in range(5):
                step()
                lun.sync()
                rc = cudart.cudaMemcpyPeer(ctypes.c_void_p(peer.data_ptr()), 1, ctypes.c_void_p(src_ptr), 0, ctypes.c_size_t(NB * 512))
                assert rc == 0
            torch.cuda.synchronize(0)
            ms2 = (time.perf_counter() - t0) / 5 * 1e3
            out["plain_then_peer_copy"] = {"ms_per_pass": ms2, "gbs": len(t.reqs) * 131072 / ms2 / 1e6}
            del peer
    c = np.frombuffer(d_cpls.cpu().numpy().tobytes(), dtype=abi.cpl_dtype)
    assert not c["status"].any()
    if label != "plain":
        a = lib.bdev_read_raw(name, 0, 1 << 20, replica=0)
        b = lib.bdev_read_raw(name, 0, 1 << 20, replica=1)
        assert (a == b).all() and a.any(), "replicas differ"
    lib.remove_vhost_scsi_target(label + ".ctl", 0)
    lib.remove_vhost_controller(label + ".ctl")
    lib.delete_bdev(name)
out["nvlink_note"] = "mirror egress = 1 x payload over NVLink 5 (900 GB/s/dir nominal, 770 GB/s measured peer copy)"
print(json.dumps(out))

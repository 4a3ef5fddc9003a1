"""Helpers shared by the parity tests: run one trace through a checker or through the CUDA path."""
from __future__ import annotations

import hashlib
import itertools

import numpy as np

from oim_b200 import abi, traces

_seq = itertools.count()


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_oracle(cls, trace: traces.Trace, num_blocks: int, *, block_size: int = 512, store_seed: int | None = 7,
               arena_seed: int = 0x5EED, target: int = 0, removed: bool = False, name: str | None = None,
               scsi_dev_id: int | None = None):
    """-> (cpls, arena, store) after replaying `trace` on a CPU checker (RefOracle / PortOracle).
    name / scsi_dev_id: the identity INQUIRY reports (bdev name, global SCSI device id)."""
    o = cls(num_blocks, block_size, target, name=name, scsi_dev_id=scsi_dev_id)
    try:
        if store_seed is not None:
            o.store[:] = traces.pattern_bytes(store_seed, 0, o.store.size)
        if removed:
            o.set_removed(target, True)
        arena = np.zeros(trace.arena_bytes, dtype=np.uint8)
        traces.fill_arena(arena, trace, arena_seed)
        cpls = o.submit(trace.reqs, trace.bind(arena.ctypes.data))
        return cpls, arena, o.store.copy()
    finally:
        o.close()


def run_cuda(lib, trace: traces.Trace, num_blocks: int, *, block_size: int = 512, store_seed: int | None = 7,
             arena_seed: int = 0x5EED, target: int = 0, mem: str = "device", removed: bool = False,
             queue_size: int = 1024, name: str | None = None):
    """Same replay through liboimgpu.so on cuda:0.  mem="device": client buffers in HBM (a torch
    tensor); mem="host": client buffers in pinned, registered host memory."""
    import torch
    tag = next(_seq)
    bname, cname = name or f"par{tag}", f"parctl{tag}"
    lib.construct_malloc_bdev(num_blocks, block_size, name=bname, device=0)
    lib.construct_vhost_scsi_controller(cname)
    lib.add_vhost_scsi_lun(cname, target, bname)
    host_arena = np.zeros(trace.arena_bytes, dtype=np.uint8)
    traces.fill_arena(host_arena, trace, arena_seed)
    try:
        if store_seed is not None:
            lib.bdev_write_raw(bname, 0, traces.pattern_bytes(store_seed, 0, num_blocks * block_size))
        with lib.Lun(cname, target, num_queues=1, queue_size=queue_size) as lun:
            if removed:
                lun.set_removed(True)
            if mem == "device":
                dev = torch.from_numpy(host_arena).to("cuda:0")
                torch.cuda.synchronize()
                cpls = lun.run(trace.reqs, trace.bind(dev.data_ptr()))
                torch.cuda.synchronize()
                arena = dev.cpu().numpy()
            else:
                pinned = torch.from_numpy(host_arena.copy()).pin_memory()
                cpls = lun.run(trace.reqs, trace.bind(pinned.data_ptr()))
                arena = pinned.numpy().copy()
            stats = lun.iostat()
        store = lib.bdev_read_raw(bname, 0, num_blocks * block_size)
        return cpls, arena, store, stats
    finally:
        lib.remove_vhost_scsi_target(cname, target)
        lib.remove_vhost_controller(cname)
        lib.delete_bdev(bname)


def assert_cpls_equal(got: np.ndarray, want: np.ndarray, reqs: np.ndarray | None = None, what: str = ""):
    assert len(got) == len(want), f"{what}: {len(got)} completions, expected {len(want)}"
    for f in abi.CPL_PARITY_FIELDS:
        a, b = got[f].reshape(len(got), -1), want[f].reshape(len(want), -1)
        bad = np.nonzero((a != b).any(axis=1))[0]
        if len(bad):
            i = int(bad[0])
            extra = f"\nreq={reqs[i]}" if reqs is not None else ""
            raise AssertionError(f"{what}: completion field '{f}' differs at {len(bad)} requests, first #{i}: "
                                 f"got {got[i]} want {want[i]}{extra}")

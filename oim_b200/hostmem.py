"""Host-side placement for the host-buffer (PCIe) paths: run the process - and therefore its pinned client
buffers, rings and bounce buffers, which are first-touched by it - on the NUMA node the GPU hangs off, and
measure what the box's host side can take when N GPUs store into host memory at once.

Why: with one process per GPU and no placement, ranks of an 8-GPU job share whatever node the scheduler put
them on; D2H payload then crosses the inter-socket link and per-GPU PCIe throughput falls (round 1: 50.9 GB/s at
N=1, 30.4 GB/s per GPU at N=8).  The reference has the same concern and solves it the same way: a vhost
controller is pinned to a reactor core by `cpumask` (S/lib/vhost/vhost.c:560-590), next to its memory."""
from __future__ import annotations

import ctypes
import os
import re
import subprocess


def gpu_pci_bdf(index: int) -> str | None:
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip()
    except (OSError, subprocess.TimeoutExpired):
        return None
    m = re.match(r"^([0-9A-Fa-f]{4,8}):([0-9A-Fa-f]{2}:[0-9A-Fa-f]{2}\.[0-9A-Fa-f])$", out)
    if not m:
        return None
    return f"{m.group(1)[-4:]}:{m.group(2)}".lower()


def gpu_numa_node(index: int) -> int | None:
    bdf = gpu_pci_bdf(index)
    if not bdf:
        return None
    try:
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def _cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def node_cpus(node: int) -> set[int]:
    try:
        return _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
    except OSError:
        return set()


def bind_to_gpu_node(index: int) -> dict:
    """Pin this process to the CPUs of the GPU's NUMA node and prefer that node's memory (MPOL_PREFERRED: a
    container whose cpuset lacks the node must still run).  Returns what was done, for the bench line."""
    info: dict = {"gpu": index, "node": None, "cpus_bound": 0, "mempolicy": "default"}
    node = gpu_numa_node(index)
    if node is None:
        info["note"] = "GPU's NUMA node unknown (no sysfs numa_node): left to the scheduler"
        return info
    info["node"] = node
    allowed = os.sched_getaffinity(0)
    cpus = node_cpus(node) & allowed
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
            info["cpus_bound"] = len(cpus)
        except OSError as e:
            info["note"] = f"sched_setaffinity: {e}"
    else:
        info["note"] = f"no CPU of node {node} in this process's cpuset ({len(allowed)} CPUs allowed)"
    import platform
    if platform.machine() != "x86_64":
        info["mempolicy"] = "default (set_mempolicy syscall number known for x86_64 only)"
        return info
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238           # x86_64
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        info["mempolicy"] = f"preferred node {node}" if rc == 0 else f"set_mempolicy failed (errno {ctypes.get_errno()})"
    except (OSError, AttributeError) as e:
        info["mempolicy"] = f"unavailable: {e}"
    return info


def concurrent_d2h_probe(torch, device: int, barrier, seconds: float = 0.4, nbytes: int = 256 << 20) -> dict:
    """every rank at once: (a) cudaMemcpyAsync device -> pinned host (the copy engine), (b) SM-originated stores into
    the same pinned buffer (what the movers do on the READ path), each for `seconds`.  GB/s of THIS rank; the caller
    sums over ranks for the box's ceiling."""
    from . import lib
    torch.cuda.set_device(device)
    dev = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{device}")
    dev.random_(0, 255)
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    out = {}
    s = torch.cuda.Stream(device=device)
    for kind in ("copy_engine", "sm_stores"):
        def one():
            if kind == "copy_engine":
                with torch.cuda.stream(s):
                    host.copy_(dev, non_blocking=True)
            else:
                L.copy(host.data_ptr(), dev.data_ptr(), nbytes)
        if kind == "sm_stores":
            name = lib.construct_malloc_bdev(2048, 512, name=f"probe{device}_{os.getpid()}", device=device)
            lib.construct_vhost_scsi_controller(f"probe.{device}.{os.getpid()}")
            lib.add_vhost_scsi_lun(f"probe.{device}.{os.getpid()}", 0, name)
            L = lib.Lun(f"probe.{device}.{os.getpid()}", 0, num_queues=1, queue_size=32)
            timer = lib.Timer()
        one()
        torch.cuda.synchronize(device)
        if kind == "sm_stores":
            L.sync()
        barrier()
        import time
        t0, n = time.perf_counter(), 0
        if kind == "copy_engine":
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s):
                a.record()
            while time.perf_counter() - t0 < seconds:
                one()
                n += 1
                s.synchronize()
            with torch.cuda.stream(s):
                b.record()
            b.synchronize()
            ms = a.elapsed_time(b)
        else:
            timer.start(L)
            while time.perf_counter() - t0 < seconds:
                one()
                n += 1
                L.sync()
            timer.stop(L)
            L.sync()
            ms = timer.elapsed_ms()
        barrier()
        out[kind + "_gbs"] = n * nbytes / (ms / 1e3) / 1e9
        if kind == "sm_stores":
            L.close()
            lib.remove_vhost_scsi_target(f"probe.{device}.{os.getpid()}", 0)
            lib.remove_vhost_controller(f"probe.{device}.{os.getpid()}")
            lib.delete_bdev(name)
    del dev, host
    return out

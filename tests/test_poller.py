"""Persistent mode: the resident reactor kernel served by doorbells (GPU only).  Every test arms the
kernel's idle watchdog so a failing assertion cannot leave a spinning kernel behind."""
import time

import numpy as np
import pytest

import util
from oim_b200 import abi, traces, vring

pytestmark = pytest.mark.gpu
WATCHDOG_MS = 4000


def _setup(gpu, name, nb):
    gpu.construct_malloc_bdev(nb, 512, name=name, device=0)
    gpu.construct_vhost_scsi_controller(name + ".ctl")
    gpu.add_vhost_scsi_lun(name + ".ctl", 0, name)
    gpu.bdev_write_raw(name, 0, traces.pattern_bytes(7, 0, nb * 512))


def _teardown(gpu, name):
    gpu.remove_vhost_scsi_target(name + ".ctl", 0)
    gpu.remove_vhost_controller(name + ".ctl")
    gpu.delete_bdev(name)


@pytest.mark.timeout(120)
def test_poller_serves_ring_doorbells(gpu, oracles):
    """closed loop, qd=32 per queue: submit 32, ring the doorbell, wait for the completion counter"""
    import torch
    nb, nq, rounds = 1 << 16, 4, 12
    t = traces.partitioned_queues(nq, 32 * rounds, nb, pattern="randrw", read_pct=60, io_blocks=8)
    host = np.zeros(t.arena_bytes, dtype=np.uint8)
    traces.fill_arena(host, t)
    with oracles.PortOracle(nb) as o:
        o.store[:] = traces.pattern_bytes(7, 0, o.store.size)
        oa = host.copy()
        want = o.submit(t.reqs, t.bind(oa.ctypes.data))
        want_store = o.store.copy()
    _setup(gpu, "pl0", nb)
    try:
        arena = torch.from_numpy(host.copy()).pin_memory()
        iovs = t.bind(arena.data_ptr())
        per_q = 32 * rounds
        got = np.zeros(len(t.reqs), dtype=abi.cpl_dtype)
        with gpu.Lun("pl0.ctl", 0, num_queues=nq, queue_size=64) as lun:
            assert lun.start_poller(idle_timeout_ms=WATCHDOG_MS) >= nq      # workers: one per queue, more when they share queues
            launches_before = None
            try:
                for r in range(rounds):
                    for q in range(nq):
                        lo = q * per_q + r * 32
                        part = t.reqs[lo:lo + 32].copy()
                        i0 = int(part["iov_start"][0])
                        part["iov_start"] -= np.uint32(i0)
                        lun.submit(q, part, iovs[i0:i0 + 32])
                    assert lun.kick() == nq                   # doorbells, not launches
                    for q in range(nq):
                        c = lun.poll(q, 32, wait=True)
                        assert len(c) == 32
                        got[q * per_q + r * 32:q * per_q + r * 32 + 32] = c
                assert lun.poller_running()
            finally:
                lun.stop_poller()
            stats = lun.iostat()
        assert stats["kernel_launches"] == 1, "one resident kernel served every round"
        util.assert_cpls_equal(got, want, t.reqs)
        assert (arena.numpy() == oa).all()
        assert (gpu.bdev_read_raw("pl0", 0, nb * 512) == want_store).all()
    finally:
        _teardown(gpu, "pl0")


@pytest.mark.timeout(120)
def test_poller_serves_guest_virtqueue(gpu, oracles):
    """a guest image in pinned memory: the 'guest' publishes avail->idx in two steps, the resident
    kernel answers through the used ring without any host call"""
    import torch
    nb = 32768
    t = traces.fuzz_trace(48, nb, seed=77, max_io_blocks=16, arena_bytes=8 << 20, include_malformed=False)
    a0 = np.zeros(t.arena_bytes, dtype=np.uint8)
    traces.fill_arena(a0, t)
    rq = vring.requests_from_trace(t, a0)
    img = vring.build_image(rq, ring_size=256, seed=3, mutate=False)
    n = img.meta["placed"]
    assert n == 48
    ref = vring.build_image(rq, ring_size=256, seed=3, mutate=False)
    with oracles.PortOracle(nb) as o:
        o.store[:] = traces.pattern_bytes(7, 0, o.store.size)
        o.vq_process(ref)
        want_store = o.store.copy()
    _setup(gpu, "pv0", nb)
    try:
        guest = torch.from_numpy(img.arena.copy()).pin_memory()
        g = guest.numpy()
        avail_idx = g[img.avail_off + 2:img.avail_off + 4].view("<u2")
        used_idx = g[img.used_off + 2:img.used_off + 4].view("<u2")
        avail_idx[0] = 0                                       # nothing published yet
        base = guest.data_ptr()
        with gpu.Lun("pv0.ctl", 0, num_queues=1, queue_size=32) as lun:
            lun.set_mem_table(img.region_table(base))
            lun.vq_attach(0, base + img.desc_off, base + img.avail_off, base + img.used_off, img.ring_size, 0, 0)
            lun.start_poller(idle_timeout_ms=WATCHDOG_MS)
            try:
                for target in (20, n):
                    avail_idx[0] = target                      # the guest's "kick" is just this store
                    deadline = time.time() + 20
                    while int(used_idx[0]) != target and time.time() < deadline:
                        time.sleep(0.0005)
                    assert int(used_idx[0]) == target, f"used->idx stuck at {int(used_idx[0])}"
            finally:
                lun.stop_poller()
            la, lu = lun.vq_detach(0)
        assert (la, lu) == (n, n)
        got = img.masked(g.copy())
        assert (got == ref.masked(ref.arena)).all()
        assert (gpu.bdev_read_raw("pv0", 0, nb * 512) == want_store).all()
    finally:
        _teardown(gpu, "pv0")


@pytest.mark.timeout(60)
def test_poller_watchdog_and_busy_errors(gpu):
    import errno
    _setup(gpu, "pw0", 4096)
    try:
        with gpu.Lun("pw0.ctl", 0, num_queues=2, queue_size=32) as lun:
            lun.start_poller(idle_timeout_ms=300)
            with pytest.raises(gpu.OimGpuError) as e:
                lun.iostat()                                   # would need to drain the stream
            assert e.value.rc == -errno.EBUSY
            t0 = time.time()
            while lun.poller_running() and time.time() - t0 < 10:
                time.sleep(0.01)
            assert not lun.poller_running(), "idle watchdog did not fire"
            lun.stop_poller()
            assert lun.iostat()["kernel_launches"] == 1
    finally:
        _teardown(gpu, "pw0")


@pytest.mark.gpu
@pytest.mark.timeout(180)
def test_poller_guest_reuses_ring_and_buffers(gpu, oracles):
    """What a real guest does all day under a resident kernel: the SAME descriptor slots, request headers,
    indirect tables and data buffers are rewritten by the CPU with different contents round after round,
    with nothing but the avail index as a signal.  Every round must see the new contents, never what
    the GPU read from those addresses in an earlier round."""
    import torch
    nb, rounds = 32768, 6
    imgs, refs = [], []
    with oracles.PortOracle(nb) as o:
        o.store[:] = traces.pattern_bytes(7, 0, o.store.size)
        la = lu = 0
        for r in range(rounds):
            t = traces.fuzz_trace(40, nb, seed=900 + r, max_io_blocks=32, arena_bytes=8 << 20, include_malformed=False)
            a0 = np.zeros(t.arena_bytes, dtype=np.uint8)
            traces.fill_arena(a0, t)
            rq = vring.requests_from_trace(t, a0)
            # same seed -> same ring slots, header / table / buffer addresses in every round; different requests
            img = vring.build_image(rq, ring_size=256, seed=3, mutate=False, pattern_seed=0x5EED + r)
            ref = vring.build_image(rq, ring_size=256, seed=3, mutate=False, pattern_seed=0x5EED + r)
            assert img.meta["placed"] == 40
            o.vq_process(ref)                                      # each image is a fresh ring for the oracle
            imgs.append(img)
            refs.append(ref)
        want_store = o.store.copy()
    _setup(gpu, "pv1", nb)
    try:
        guest = torch.from_numpy(imgs[0].arena.copy()).pin_memory()
        g = guest.numpy()
        img0 = imgs[0]
        avail_idx = g[img0.avail_off + 2:img0.avail_off + 4].view("<u2")
        used_idx = g[img0.used_off + 2:img0.used_off + 4].view("<u2")
        avail_idx[0] = 0
        base = guest.data_ptr()
        with gpu.Lun("pv1.ctl", 0, num_queues=1, queue_size=32) as lun:
            lun.set_mem_table(img0.region_table(base))
            lun.vq_attach(0, base + img0.desc_off, base + img0.avail_off, base + img0.used_off, img0.ring_size, 0, 0)
            lun.start_poller(idle_timeout_ms=WATCHDOG_MS)
            try:
                done = 0
                for r in range(rounds):
                    img = imgs[r]
                    if r:
                        # the guest rewrites its memory in place: everything except the ring indices, then
                        # the avail ring entries at the positions the next 40 heads go to
                        keep_a, keep_u = int(avail_idx[0]), int(used_idx[0])
                        lo, hi = img.avail_off, img.used_off + 4 + 8 * img.ring_size
                        g[:lo] = img.arena[:lo]
                        g[hi:] = img.arena[hi:]
                        ring = g[img.avail_off + 4:img.avail_off + 4 + 2 * img.ring_size].view("<u2")
                        src = img.arena[img.avail_off + 4:img.avail_off + 4 + 2 * img.ring_size].view("<u2")
                        for k in range(40):
                            ring[(keep_a + k) % img.ring_size] = src[k]
                        assert (int(avail_idx[0]), int(used_idx[0])) == (keep_a, keep_u)
                    done += 40
                    avail_idx[0] = done
                    deadline = time.time() + 20
                    while int(used_idx[0]) != done and time.time() < deadline:
                        time.sleep(0.0005)
                    assert int(used_idx[0]) == done, f"round {r}: used->idx stuck at {int(used_idx[0])}"
                    # everything outside the two rings equals what the oracle left for this round's image
                    want = refs[r].arena
                    lo, hi = img.avail_off, img.used_off + 4 + 8 * img.ring_size
                    assert (g[:lo] == want[:lo]).all(), f"round {r}: guest memory differs at {np.nonzero(g[:lo] != want[:lo])[0][:8]}"
                    assert (g[hi:] == want[hi:]).all(), f"round {r}: guest memory differs at {hi + np.nonzero(g[hi:] != want[hi:])[0][:8]}"
            finally:
                lun.stop_poller()
        assert (gpu.bdev_read_raw("pv1", 0, nb * 512) == want_store).all()
    finally:
        _teardown(gpu, "pv1")


@pytest.mark.gpu
@pytest.mark.timeout(120)
@pytest.mark.parametrize("resident", [False, True])
def test_slot_ring_passes_that_straddle_the_ring_end(gpu, oracles, resident):
    """Submissions of uneven size on one 64-slot ring: passes start at slot 16, 48 (-> wraps to 16), ... so a pass's
    slots lie in two pieces of the ring - the TMA unit fetches them with two bulk copies (lun_kernel.cu issue_fetch).
    Launch per kick and resident poller; results equal the oracle's sequential execution."""
    import torch
    nb = 1 << 15
    sizes = [16, 32, 32, 7, 32, 25, 32, 32, 1, 31, 32, 32]
    n = sum(sizes)
    t = traces.uniform_trace(n, nb, io_blocks=8, pattern="randrw", read_pct=50, seed=91, lba_span=4096)
    host = np.zeros(t.arena_bytes, dtype=np.uint8)
    traces.fill_arena(host, t)
    with oracles.PortOracle(nb) as o:
        o.store[:] = traces.pattern_bytes(7, 0, o.store.size)
        oa = host.copy()
        want = o.submit(t.reqs, t.bind(oa.ctypes.data))
        want_store = o.store.copy()
    name = "wrap1" if resident else "wrap0"
    _setup(gpu, name, nb)
    try:
        arena = torch.from_numpy(host.copy()).pin_memory()
        iovs = t.bind(arena.data_ptr())
        got = []
        with gpu.Lun(name + ".ctl", 0, num_queues=1, queue_size=64) as lun:
            if resident:
                lun.start_poller(idle_timeout_ms=WATCHDOG_MS)
            try:
                lo = 0
                for k in sizes:
                    part = t.reqs[lo:lo + k].copy()
                    i0 = int(part["iov_start"][0])
                    part["iov_start"] -= np.uint32(i0)
                    lun.submit(0, part, iovs[i0:i0 + k])
                    lun.kick()
                    c = lun.poll(0, k, wait=True)
                    assert len(c) == k
                    got.append(c)
                    lo += k
            finally:
                if resident:
                    lun.stop_poller()
        util.assert_cpls_equal(np.concatenate(got), want, t.reqs)
        assert (arena.numpy() == oa).all()
        assert (gpu.bdev_read_raw(name, 0, nb * 512) == want_store).all()
    finally:
        _teardown(gpu, name)

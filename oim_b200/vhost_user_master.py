"""A vhost-user MASTER (what QEMU's vhost-user-scsi-pci is to the slave), for the transport tests.

Speaks the protocol S/lib/vhost/rte_vhost/vhost_user.{h,c} implements: 12-byte header
{u32 request, u32 flags, u32 size} + payload, file descriptors as SCM_RIGHTS.  Guest memory is a memfd
shared with the slave; rings and buffers are laid out in it by oim_b200.vring.

Used by the tests (against oim-gpu-vhost and against the reference's own transport
oracle/_ref/liboim_ref_vhost.so, with the same script) and by bench.py's vhost-user leg."""
from __future__ import annotations

import mmap
import os
import select
import socket
import struct
import time

import numpy as np

# VhostUserRequest (vhost_user.h:65-96)
GET_FEATURES, SET_FEATURES, SET_OWNER, RESET_OWNER, SET_MEM_TABLE = 1, 2, 3, 4, 5
SET_LOG_BASE, SET_LOG_FD, SET_VRING_NUM, SET_VRING_ADDR, SET_VRING_BASE, GET_VRING_BASE = 6, 7, 8, 9, 10, 11
SET_VRING_KICK, SET_VRING_CALL, SET_VRING_ERR, GET_PROTOCOL_FEATURES, SET_PROTOCOL_FEATURES = 12, 13, 14, 15, 16
GET_QUEUE_NUM, SET_VRING_ENABLE, GET_CONFIG, SET_CONFIG = 17, 18, 24, 25
VERSION, REPLY_MASK, NEED_REPLY = 0x1, 0x4, 0x8
NOFD_MASK = 0x100

F_LOG_ALL, F_INDIRECT_DESC, F_EVENT_IDX, F_PROTOCOL_FEATURES, F_VERSION_1 = 26, 28, 29, 30, 32
SCSI_F_INOUT, SCSI_F_HOTPLUG, SCSI_F_CHANGE = 0, 1, 2
PF_MQ, PF_LOG_SHMFD, PF_RARP, PF_REPLY_ACK, PF_NET_MTU, PF_CONFIG = 0, 1, 2, 3, 4, 9

UVA_BASE = 0x7F00_0000_0000       # where "QEMU" pretends guest RAM is mapped in its own address space


class GuestRam:
    """one memfd = the guest's RAM; regions are windows of it"""

    def __init__(self, nbytes: int):
        assert nbytes % (2 << 20) == 0, "the reference refuses regions that are not 2 MiB multiples (vhost.c:1091-1097)"
        self.fd = os.memfd_create("guest-ram")
        os.ftruncate(self.fd, nbytes)
        self.map = mmap.mmap(self.fd, nbytes)
        self.mem = np.frombuffer(self.map, dtype=np.uint8)
        self.size = nbytes

    def close(self):
        self.mem = None
        try:
            self.map.close()
        except BufferError:
            pass
        os.close(self.fd)


class Master:
    def __init__(self, path: str, timeout: float = 10.0):
        self.s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        t0 = time.time()
        while True:
            try:
                self.s.connect(path)
                break
            except (FileNotFoundError, ConnectionRefusedError):
                if time.time() - t0 > timeout:
                    raise
                time.sleep(0.01)
        self.s.settimeout(timeout)
        self.log: list[tuple] = []          # (request, reply payload or None): the transcript compared between slaves

    # ---- wire ----
    def send(self, req: int, payload: bytes = b"", fds=(), need_reply: bool = False) -> None:
        flags = VERSION | (NEED_REPLY if need_reply else 0)
        msg = struct.pack("<III", req, flags, len(payload)) + payload
        if fds:
            socket.send_fds(self.s, [msg], list(fds))
        else:
            self.s.sendall(msg)

    def recv(self) -> tuple[int, int, bytes]:
        hdr = b""
        while len(hdr) < 12:
            chunk = self.s.recv(12 - len(hdr))
            if not chunk:
                raise ConnectionError("slave closed the connection")
            hdr += chunk
        req, flags, size = struct.unpack("<III", hdr)
        body = b""
        while len(body) < size:
            chunk = self.s.recv(size - len(body))
            if not chunk:
                raise ConnectionError("slave closed the connection")
            body += chunk
        return req, flags, body

    def get_u64(self, req: int) -> int:
        self.send(req)
        r, flags, body = self.recv()
        assert r == req and (flags & REPLY_MASK) and (flags & 3) == VERSION, (r, flags)
        v = struct.unpack("<Q", body[:8])[0]
        self.log.append((req, v))
        return v

    def set_u64(self, req: int, v: int, fds=(), need_reply: bool = False):
        self.send(req, struct.pack("<Q", v), fds, need_reply)
        ack = None
        if need_reply:
            r, flags, body = self.recv()
            assert r == req and (flags & REPLY_MASK)
            ack = struct.unpack("<Q", body[:8])[0]
        self.log.append((req, ack))
        return ack

    # ---- messages ----
    def set_mem_table(self, regions, need_reply: bool = False):
        """regions: [(guest_phys_addr, size, userspace_addr, mmap_offset, fd)]"""
        body = struct.pack("<II", len(regions), 0)
        for g, s, u, o, _ in regions:
            body += struct.pack("<QQQQ", g, s, u, o)
        self.send(SET_MEM_TABLE, body, [r[4] for r in regions], need_reply)
        ack = None
        if need_reply:
            ack = struct.unpack("<Q", self.recv()[2][:8])[0]
        self.log.append((SET_MEM_TABLE, ack))
        return ack

    def vring_state(self, req: int, index: int, num: int):
        self.send(req, struct.pack("<II", index, num))
        self.log.append((req, None))

    def set_vring_addr(self, index: int, desc: int, used: int, avail: int, log: int = 0, flags: int = 0):
        self.send(SET_VRING_ADDR, struct.pack("<IIQQQQ", index, flags, desc, used, avail, log))
        self.log.append((SET_VRING_ADDR, None))

    def get_vring_base(self, index: int) -> int:
        self.send(GET_VRING_BASE, struct.pack("<II", index, 0))
        r, flags, body = self.recv()
        assert r == GET_VRING_BASE and len(body) == 8
        idx, num = struct.unpack("<II", body)
        assert idx == index
        self.log.append((GET_VRING_BASE, num))
        return num

    def get_config(self, size: int = 36) -> tuple[int, bytes]:
        self.send(GET_CONFIG, struct.pack("<III", 0, size, 0) + bytes(size))
        r, flags, body = self.recv()
        self.log.append((GET_CONFIG, len(body)))
        return len(body), body

    def close(self):
        self.s.close()


class Queue:
    """one virtqueue of the session: ring addresses inside guest RAM + its two eventfds"""

    def __init__(self, index: int, size: int, desc_off: int, avail_off: int, used_off: int):
        self.index, self.size = index, size
        self.desc_off, self.avail_off, self.used_off = desc_off, avail_off, used_off
        self.kick = os.eventfd(0, os.EFD_NONBLOCK)
        self.call = os.eventfd(0, os.EFD_NONBLOCK)

    def setup(self, m: Master, base: int = 0):
        m.vring_state(SET_VRING_NUM, self.index, self.size)
        m.vring_state(SET_VRING_BASE, self.index, base)
        m.set_vring_addr(self.index, UVA_BASE + self.desc_off, UVA_BASE + self.used_off, UVA_BASE + self.avail_off)
        m.set_u64(SET_VRING_KICK, self.index, [self.kick])
        m.set_u64(SET_VRING_CALL, self.index, [self.call])

    def notify(self):
        os.eventfd_write(self.kick, 1)

    def drain_calls(self) -> int:
        try:
            return os.eventfd_read(self.call)
        except BlockingIOError:
            return 0

    def wait_used(self, ram: GuestRam, want_idx: int, timeout: float = 20.0) -> int:
        """block on the call eventfd until used->idx reaches want_idx; -> number of interrupts seen"""
        t0, calls = time.time(), 0
        while True:
            idx = int(ram.mem[self.used_off + 2:self.used_off + 4].view("<u2")[0])
            if idx == want_idx & 0xFFFF:
                return calls + self.drain_calls()
            if time.time() - t0 > timeout:
                raise TimeoutError(f"queue {self.index}: used idx {idx}, wanted {want_idx}")
            r, _, _ = select.select([self.call], [], [], 0.05)
            if r:
                calls += self.drain_calls()

    def close(self):
        os.close(self.kick)
        os.close(self.call)

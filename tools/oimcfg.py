#!/usr/bin/env python
"""save / load the daemon's configuration over its JSON-RPC socket, the way SPDK's scripts/rpc.py save_config /
load_config do for the reference: `save` asks every subsystem for the calls that rebuild its state
(get_subsystems + get_subsystem_config, S/lib/event/rpc/subsystem_rpc.c:40-129) and prints
{"subsystems": [{"subsystem": ..., "config": [{"method", "params"}, ...]}, ...]}; `load` replays them in order.
  python tools/oimcfg.py save /var/tmp/spdk.sock > cfg.json
  python tools/oimcfg.py load /var/tmp/spdk.sock < cfg.json
set_bdev_options is a start-up option of SPDK's bdev_io pools (no such pools here): it is saved for fidelity and
skipped on load."""
import json
import socket
import sys


class Rpc:
    def __init__(self, path):
        self.s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self.s.connect(path)
        self.id = 0

    def call(self, method, params=None):
        self.id += 1
        req = {"jsonrpc": "2.0", "method": method, "id": self.id}
        if params is not None:
            req["params"] = params
        self.s.sendall((json.dumps(req) + "\n").encode())
        buf = b""
        while not buf.endswith(b"\n"):
            chunk = self.s.recv(1 << 20)
            if not chunk:
                raise ConnectionError("daemon closed the connection")
            buf += chunk
        rep = json.loads(buf)
        if "error" in rep:
            raise RuntimeError(f"{method}: code: {rep['error']['code']} msg: {rep['error']['message']}")
        return rep["result"]


def save(rpc: Rpc) -> dict:
    return {"subsystems": [{"subsystem": s["subsystem"], "config": rpc.call("get_subsystem_config", {"name": s["subsystem"]})}
                           for s in rpc.call("get_subsystems")]}


def load(rpc: Rpc, cfg: dict) -> int:
    n = 0
    for sub in cfg["subsystems"]:
        for item in sub["config"] or []:
            if item["method"] == "set_bdev_options":
                continue
            rpc.call(item["method"], item.get("params"))
            n += 1
    return n


if __name__ == "__main__":
    if len(sys.argv) != 3 or sys.argv[1] not in ("save", "load"):
        sys.exit(__doc__)
    r = Rpc(sys.argv[2])
    if sys.argv[1] == "save":
        json.dump(save(r), sys.stdout, indent=2)
        print()
    else:
        print(f"{load(r, json.load(sys.stdin))} calls replayed", file=sys.stderr)

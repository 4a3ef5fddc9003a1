"""Regenerate tests/golden/*.npz from the REFERENCE (run in the container that has /root/reference).

    python tests/golden/make_golden.py

Each fixture is one seeded fuzz trace replayed through oracle/_ref/liboim_ref.so (the reference's own
SPDK sources, see oracle/Makefile): the request array, the SG table (arena offsets), the parameter
payloads, the completions the reference produced and SHA-256 digests of the final client arena
and backing store.  Tests replay the same inputs through the C restatement (CPU) and the CUDA path
(GPU box, where /root/reference does not exist) and compare bit for bit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oim_b200 import traces  # noqa: E402
from oracle import bindings  # noqa: E402
import util  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NUM_BLOCKS = 16384          # 8 MiB disk
CASES = [  # (name, seed, n, kwargs)
    ("fuzz_a", 101, 320, {}),
    ("fuzz_b", 102, 320, {}),
    ("fuzz_c", 103, 320, {"max_io_blocks": 8}),
    ("fuzz_big", 104, 96, {"max_io_blocks": 1024, "arena_bytes": 24 << 20}),
    ("fuzz_clean", 105, 400, {"include_malformed": False}),
    ("fuzz_removed", 106, 120, {}),
    # SG lists cut from ONE client buffer: elements continue each other (the CUDA path moves such runs as one segment)
    ("fuzz_runs", 107, 320, {"contiguous": True}),
    ("fuzz_runs_big", 108, 96, {"max_io_blocks": 1024, "arena_bytes": 24 << 20, "contiguous": True}),
]


def main():
    bindings.build()
    only = set(sys.argv[1:])        # `make_golden.py fuzz_runs ...` regenerates just those
    for name, seed, n, kw in CASES:
        if only and name not in only:
            continue
        t = traces.fuzz_trace(n, NUM_BLOCKS, seed=seed, **kw)
        removed = name == "fuzz_removed"
        cpls, arena, store = util.run_oracle(bindings.RefOracle, t, NUM_BLOCKS, removed=removed)
        pay = t.meta["param_payloads"]
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            reqs=t.reqs, iovs=t.iovs, arena_bytes=np.int64(t.arena_bytes), num_blocks=np.int64(NUM_BLOCKS),
            removed=np.bool_(removed), cpls=cpls,
            pay_off=np.array([o for o, _ in pay], dtype=np.int64),
            pay_len=np.array([len(d) for _, d in pay], dtype=np.int64),
            pay_data=np.concatenate([d for _, d in pay]) if pay else np.zeros(0, np.uint8),
            arena_sha=np.array(util.sha(arena)), store_sha=np.array(util.sha(store)),
            source=np.array(bindings.RefOracle.describe()))
        print(name, len(t), "status hist", dict(zip(*np.unique(cpls["status"], return_counts=True))))


def primary():
    """SPC primary commands (INQUIRY / MODE SENSE / REPORT LUNS / MODE SELECT), bdev named Malloc0"""
    n, seed = 400, 909
    t = traces.primary_trace(n, seed=seed)
    with bindings.RefOracle(16384, name="Malloc0") as probe:
        dev_id = probe.scsi_dev_id
    cpls, arena, _ = util.run_oracle(bindings.RefOracle, t, 16384, name="Malloc0")
    np.savez_compressed(os.path.join(HERE, "primary.npz"), n=np.int64(n), seed=np.int64(seed), reqs=t.reqs,
                        cpls=cpls, arena_sha=np.array(util.sha(arena)), scsi_dev_id=np.int64(dev_id),
                        source=np.array(bindings.RefOracle.describe()))
    print("primary", n, "scsi_dev_id", dev_id, dict(zip(*np.unique(cpls["status"], return_counts=True))))


if __name__ == "__main__":
    if len(sys.argv) == 1:
        primary()
    main()

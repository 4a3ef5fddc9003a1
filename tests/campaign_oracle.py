"""Offline differential campaign: the C restatement against the compiled reference on many more seeds than the
test-suite runs (CPU only; needs oracle/_ref).  python tests/campaign_oracle.py  ->  mismatch count per family."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
N = float(os.environ.get("CAMPAIGN_SCALE", "1"))      # CAMPAIGN_SCALE=0.02 for a smoke run
import numpy as np
from oracle import bindings
from oim_b200 import traces, vring, abi
import util
bad=0; t0=time.time()
# 1. fuzz traces, varied geometry
for seed in range(2000, 2000 + max(1, int(250 * N))):
    nb=[32768, 8192, 65536, 16384][seed%4]; bs=[512,512,4096,520][seed%4] if seed%7==0 else 512
    try:
        t=traces.fuzz_trace(400, nb, block_size=bs, seed=seed, max_io_blocks=[8,64,300,1024][seed%4], arena_bytes=(48<<20) if seed%4==3 else (16<<20))
    except TypeError:
        t=traces.fuzz_trace(400, nb, seed=seed, max_io_blocks=[8,64,300,1024][seed%4], arena_bytes=(48<<20) if seed%4==3 else (16<<20)); bs=512
    a=util.run_oracle(bindings.RefOracle, t, nb, block_size=bs)
    b=util.run_oracle(bindings.PortOracle, t, nb, block_size=bs)
    try:
        util.assert_cpls_equal(b[0], a[0], t.reqs, f"seed {seed}")
        assert (a[1]==b[1]).all() and (a[2]==b[2]).all()
    except AssertionError as e:
        bad+=1; print("MISMATCH fuzz", seed, str(e)[:300], flush=True)
print("fuzz done", time.time()-t0, "bad", bad, flush=True)
# 1b. the same with SG lists cut from ONE client buffer (elements continue each other: the runs the CUDA parser joins)
for seed in range(2500, 2500 + max(1, int(150 * N))):
    nb=[32768, 8192, 65536, 16384][seed%4]
    t=traces.fuzz_trace(400, nb, seed=seed, max_io_blocks=[8,64,300,1024][seed%4], arena_bytes=(48<<20) if seed%4==3 else (16<<20), contiguous=True)
    a=util.run_oracle(bindings.RefOracle, t, nb)
    b=util.run_oracle(bindings.PortOracle, t, nb)
    try:
        util.assert_cpls_equal(b[0], a[0], t.reqs, f"seed {seed}")
        assert (a[1]==b[1]).all() and (a[2]==b[2]).all()
    except AssertionError as e:
        bad+=1; print("MISMATCH fuzz-runs", seed, str(e)[:300], flush=True)
print("fuzz-runs done", time.time()-t0, "bad", bad, flush=True)
# 2. primary commands
for seed in range(3000, 3000 + max(1, int(100 * N))):
    t=traces.primary_trace(300, seed=seed)
    a=util.run_oracle(bindings.RefOracle, t, 32768, name=f"camp{seed}")
    ro=None
    b=util.run_oracle(bindings.PortOracle, t, 32768, name=f"camp{seed}", scsi_dev_id=0)
    # scsi id differs (global slot) -> skip payload compare when ids differ; compare statuses only
    try:
        util.assert_cpls_equal(b[0], a[0], t.reqs, f"primary {seed}")
    except AssertionError as e:
        bad+=1; print("MISMATCH primary", seed, str(e)[:300], flush=True)
print("primary done", time.time()-t0, "bad", bad, flush=True)

# ---- virtqueue images (with malformed chains) and multi-target controllers ----
import test_vring as TV, test_multi_target as TM
class O: RefOracle=bindings.RefOracle; PortOracle=bindings.PortOracle
for seed in range(5000, 5000 + max(1, int(150 * N))):
    nb, rq = 32768, TV.make_requests(seed)
    ring = [64, 256, 1024][seed % 3]
    want, ws = TV.run_kicks_oracle(bindings.RefOracle, rq, nb, ring, seed)
    got, gs = TV.run_kicks_oracle(bindings.PortOracle, rq, nb, ring, seed)
    ok = len(got)==len(want) and all(g[2]==w[2] and g[1]==w[1] and (g[0]==w[0]).all() for g,w in zip(got,want)) and (gs==ws).all()
    if not ok: bad+=1; print("MISMATCH vring", seed, flush=True)
print("vring done", time.time()-t0, "bad", bad, flush=True)
for seed in range(5500, 5500 + max(1, int(80 * N))):     # data descriptors that continue each other in guest memory
    nb, rq = 32768, TV.make_requests(seed, contiguous=True)
    ring = [64, 256, 1024][seed % 3]
    want, ws = TV.run_kicks_oracle(bindings.RefOracle, rq, nb, ring, seed, contiguous=True)
    got, gs = TV.run_kicks_oracle(bindings.PortOracle, rq, nb, ring, seed, contiguous=True)
    ok = len(got)==len(want) and all(g[2]==w[2] and g[1]==w[1] and (g[0]==w[0]).all() for g,w in zip(got,want)) and (gs==ws).all()
    if not ok: bad+=1; print("MISMATCH vring-runs", seed, flush=True)
print("vring-runs done", time.time()-t0, "bad", bad, flush=True)
for seed in range(6000, 6000 + max(1, int(60 * N))):
    for kind in ("fuzz","primary"):
        t = TM.multi_trace(seed, kind=kind)
        names = {1: f"c{seed}a", 3: f"c{seed}b", 6: f"c{seed}c"}
        wc, wa, ws, sid = TM.run_oracle(bindings.RefOracle, t, names)
        gc, ga, gs, _ = TM.run_oracle(bindings.PortOracle, t, names, sid)
        try:
            util.assert_cpls_equal(gc, wc, t.reqs, f"mt {seed} {kind}")
            assert (ga==wa).all() and all((gs[k]==ws[k]).all() for k in TM.TARGETS)
        except AssertionError as e:
            bad+=1; print("MISMATCH multi", seed, kind, str(e)[:200], flush=True)
print("multi done", time.time()-t0, "bad", bad, flush=True)

"""Offline differential campaign for the vhost-user slave (daemon with --control-only) against the reference's own
transport: random protocol message sequences and random control-queue requests, far more seeds than the tests run.
CPU only; needs oracle/_ref/liboim_ref_vhost.so.  python tests/campaign_transport.py"""
import os, sys, pathlib, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_vhost_user as T
from oim_b200 import vhost_user_master as vu
tmp=pathlib.Path(tempfile.mkdtemp())
ours=T.Slave("ours",tmp,["--control-only"]); ref=T.Slave("ref",tmp)
bad=0; t0=time.time()
try:
    T.provision(ours); T.provision(ref)
    for seed in range(1000, 1600):
        logs=[]
        for s in (ours, ref):
            ram=vu.GuestRam(4<<20); m=vu.Master(s.sock("scsi0"))
            try:
                T.fuzz_script(m, ram, np.random.default_rng(seed), nmsg=80); logs.append(m.log)
            except (ConnectionError, OSError, TimeoutError) as e:
                logs.append(("dropped", type(e).__name__, len(m.log)))
            finally:
                m.close(); ram.close()
            if s.p.poll() is not None:
                print("SERVER DIED", s.kind, seed, flush=True); raise SystemExit(1)
        if logs[0]!=logs[1]:
            bad+=1; print("MISMATCH", seed, flush=True)
            for a,b in zip(logs[0],logs[1]):
                if a!=b: print("  first diff", a, b); break
    print("vhost-user fuzz done", time.time()-t0, "bad", bad, flush=True)
    for seed in range(100, 140):
        a,ua=T.control_fuzz_script(ours, seed); b,ub=T.control_fuzz_script(ref, seed)
        if ua!=ub or not (a==b).all(): bad+=1; print("MISMATCH ctrl", seed, flush=True)
    print("control queue fuzz done", time.time()-t0, "bad", bad, flush=True)
finally:
    ours.close(); ref.close()
    log = open(ours.dir / "log.txt", errors="replace").read()
    hits = [ln for ln in log.splitlines() if "runtime error" in ln or "AddressSanitizer" in ln]
    print("sanitizer reports in the daemon log:", len(hits))      # meaningful with OIM_DAEMON_PATH=<asan build>
    for ln in hits[:10]:
        print("  ", ln[:200])

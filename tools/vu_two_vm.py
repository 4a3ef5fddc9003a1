"""diagnostics: two concurrent vhost-user sessions against the daemon in --poller mode"""
import sys, pathlib, tempfile, time, faulthandler
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
faulthandler.dump_traceback_later(50, exit=True)
import os
os.environ["OIM_VU_DEBUG"]="1"
import test_vhost_user as T
tmp = pathlib.Path(tempfile.mkdtemp())
s = T.Slave("ours", tmp, ["--poller"])
try:
    T.provision(s)
    a = T.Vm(s, T.make_requests(21, 64), 3); print("A connected", flush=True)
    b = T.Vm(s, T.make_requests(22, 64), 4); print("B connected", flush=True)
    try:
        b.io(); print("B io done", flush=True)
    except Exception as e:
        print("B io failed", e, flush=True)
    b.close(graceful=False); print("B closed", flush=True)
    try:
        a.io(); print("A io done", flush=True)
    except Exception as e:
        print("A io failed", e, flush=True)
    print("A base", a.close(), flush=True)
    c = T.Vm(s, T.make_requests(23, 64), 5); print("C connected", c.img.meta["placed"], flush=True)
    try:
        c.io(); print("C io done", flush=True)
    except Exception as e:
        print("C io failed", e, flush=True)
finally:
    s.close()
    print(open(tmp / "ours" / "log.txt").read()[-3500:])

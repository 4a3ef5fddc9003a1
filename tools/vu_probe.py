"""per-round timing of the vhost-user leg in both modes (diagnostics)"""
import sys, types, time, json
sys.path.insert(0, '.')
import bench
a = types.SimpleNamespace(steps=6, warmup=2)
for mode in ("kick", "poller"):
    t0 = time.time()
    r = bench.vhost_user_leg(a, 0, mode)
    print(mode, round(r["value"] / 1e6, 2), "M IOPS", round(r["ms_per_round"], 3), "ms/round", "total", round(time.time() - t0, 1), "s", flush=True)

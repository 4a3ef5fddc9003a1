"""Diagnose the one-controller-N-GPUs vhost-user path: the same master script against variants of the daemon.
gpurun --gpus 2 -- python tools/vu_multi_probe.py"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

a = argparse.Namespace(steps=6, warmup=3)
import torch  # noqa: E402
n = torch.cuda.device_count()
out = {}
only = set(os.environ.get("VU_ONLY", "").split(",")) - {""}
for label, kw in (("1gpu_1lun", dict()),
                  ("spread", dict(gpus=list(range(n)))),
                  ("spread_no_shared_queues", dict(gpus=list(range(n)), env={"OIMGPU_NO_SHARED_QUEUES": "1"})),
                  ("no_spread_p2p", dict(gpus=list(range(n)), daemon_args=["--no-spread"])),
                  ("spread_poller", dict(gpus=list(range(n)), mode="poller"))):
    if only and label not in only:
        continue
    mode = kw.pop("mode", "kick")
    try:
        r = bench.vhost_user_leg(a, 0, mode, (254,), per_q=int(os.environ.get("VU_PER_Q", 1024)), indirect=True, **kw)
        out[label] = {"miops": round(r["value"] / 1e6, 2), "ms_per_round": round(r["ms_per_round"], 2)}
    except Exception as e:  # noqa: BLE001
        out[label] = {"error": f"{type(e).__name__}: {e}"[:200]}
    print(label, out[label], flush=True)
print(json.dumps(out))

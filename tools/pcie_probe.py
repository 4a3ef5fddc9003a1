"""PCIe ceiling of the box: pinned D2H / H2D cudaMemcpy and the e2e leg's device time (evidence for
the e2e numbers in bench.py; prints one JSON line)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
out = {}
for name, (dst, src) in {"d2h": (h, d), "h2d": (d, h)}.items():
    best = 0
    for _ in range(5):
        torch.cuda.synchronize()
        t = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n / (time.perf_counter() - t) / 1e9)
    out[name + "_gbs"] = round(best, 2)
print(json.dumps(out))

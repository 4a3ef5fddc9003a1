"""Build kernel tuning variants in-tree (oim_b200/variants/, git-ignored .so files travel to the GPU
box) and, with --run, bench each one on the GPU: prints one line per variant."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oim_b200 import build  # noqa: E402

VARIANTS = [(mb, mv, st) for mb, mv, st in [(2, 7, 3), (3, 7, 3), (2, 8, 3), (3, 6, 3), (4, 5, 3), (2, 11, 3),
                                            (3, 7, 2), (3, 7, 4), (2, 15, 3)]]
VDIR = os.path.join(ROOT, "oim_b200", "variants")


def path(v):
    return os.path.join(VDIR, "liboimgpu_mb%d_mv%d_st%d.so" % v)


if "--run" not in sys.argv:
    os.makedirs(VDIR, exist_ok=True)
    for v in VARIANTS:
        build.build(force=True, defs="-DOIM_MIN_BLOCKS=%d -DOIM_MOVERS=%d -DOIM_STAGES=%d" % v, out=path(v))
        print("built", path(v))
else:
    extra = [a for a in sys.argv[1:] if a != "--run"]
    for v in VARIANTS:
        if not os.path.exists(path(v)):
            continue
        env = dict(os.environ, OIM_LIB_PATH=path(v))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "3",
                              "--no-cpu", "--no-e2e", *extra], env=env, capture_output=True, text=True)
        try:
            j = json.loads(out.stdout.strip().splitlines()[-1])
            seq = j.get("seq128k") or {}
            print("mb%d mv%d st%d" % v, "rand4k %.1f M IOPS frac %.3f" % (j["value"] / 1e6, j["roofline"]["frac"]),
                  "| seq128k %.0f GB/s frac %.3f" % (seq.get("value", 0), seq.get("hbm_frac", 0)), flush=True)
        except Exception as e:  # noqa: BLE001
            print("mb%d mv%d st%d" % v, "FAILED", e, out.stderr[-500:], flush=True)

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list for profiles/.

usage: tools/launch_summary.py launches.csv "command that was profiled" > profiles/rN_launches.md
"""
import csv
import sys
from collections import defaultdict


def main():
    path, cmd = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    per = defaultdict(list)
    ours = []
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        ns = float(r[vi].replace(",", ""))
        per[r[ki]].append(ns)
        if "oim_" in r[ki]:
            ours.append((r[ki].split("(")[0], int(ns), r[h.index("Grid Size")], r[h.index("Block Size")]))
    total = sum(sum(v) for v in per.values())
    print(f"# ncu launch list (gpu__time_duration.sum, --clock-control none) of: {cmd}")
    print("# per-launch times are cold-cache and serialised: compare shares, not absolutes")
    print("# (at:: kernels, when listed, are torch filling the synthetic store / client arenas BEFORE the timed regions;)")
    print("# a timed step is exactly one launch of a queue kernel: oim_lun_queue_kernel (one CTA per queue, slot rings),")
    print("# oim_lun_shared_queue_kernel (CTAs share queues: fewer than ~160 queues), oim_lun_vring_kernel (guest virtqueues)\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {k[:90]} | {len(v)} | {sum(v) / 1e6:.3f} | {100 * sum(v) / total:.1f}% |")
    print("\n## our kernels, in launch order\n\n| kernel | ns | grid | block |\n|---|---|---|---|")
    for k, ns, g, b in ours:
        if "oim_copy_kernel" in k or "oim_fill_kernel" in k:
            continue
        print(f"| {k} | {ns} | {g} | {b} |")


if __name__ == "__main__":
    main()

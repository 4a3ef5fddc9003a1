#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines.

usage: tools/stall_by_line.py report.ncu-rep kernel_mangled_substring [top]
Joins `ncu --page source --csv` (SASS + samples) with `nvdisasm -g` line info of the built library.
"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def line_map(kernel: str):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "oim_b200" / "liboimgpu.so")], cwd=tmp, check=True,
                   capture_output=True)
    out = subprocess.run(["nvdisasm", "-g", "-c", "lun_kernel.sm_100a.cubin"], cwd=tmp, check=True,
                         capture_output=True, text=True).stdout
    cur, inside, m = None, False, {}
    for ln in out.splitlines():
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        f = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if f:
            cur = (Path(f.group(1)).name, int(f.group(2)))
            continue
        a = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if a:
            m[int(a.group(1), 16)] = (cur, a.group(2).strip())
    return m


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    lm = line_map(kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    si, ai, src = h.index("# Samples"), 0, h.index("Source")
    base = None
    per_line, per_inst = defaultdict(int), []
    for r in rows[hi + 1:]:
        if len(r) <= si:
            continue
        addr = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
        if base is None:
            base = addr
        n = int(r[si] or 0)
        loc, _ = lm.get(addr - base, (None, None))
        per_line[loc] += n
        per_inst.append((n, addr - base, loc, r[src]))
    tot = sum(per_line.values())
    print(f"total samples {tot}")
    srcs = {}
    for (loc, n) in sorted(per_line.items(), key=lambda kv: -kv[1])[:top]:
        text = ""
        if loc:
            f = ROOT / "oim_b200" / "csrc" / loc[0]
            if f.exists():
                srcs.setdefault(f, f.read_text().splitlines())
                text = srcs[f][loc[1] - 1].strip()[:100]
        print(f"{100 * n / tot:5.1f}%  {loc}  {text}")
    print("\ntop instructions")
    for n, a, loc, s in sorted(per_inst, reverse=True)[:25]:
        print(f"{100 * n / tot:5.1f}%  {a:#06x} {loc} {s[:80]}")


if __name__ == "__main__":
    main()

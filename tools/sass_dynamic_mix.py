#!/usr/bin/env python
"""usage: sass_dynamic_mix.py <report.ncu-rep> <warp_tiles> [top]
Executed SASS instructions by opcode, per 32-env warp-tile, from the source page of an `ncu --set full --import-source on`
report (read here on the CPU box with `ncu -i`).  warp_tiles = N / 32 (x steps for the fused rollout kernel)."""
import collections
import csv
import io
import re
import subprocess
import sys

path, tiles = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
ix = {h: i for i, h in enumerate(rows[hdr])}
hist, samp, total, tot_s = collections.Counter(), collections.Counter(), 0.0, 0.0
for r in rows[hdr + 1:]:
    try:
        n, s = float(r[ix["Instructions Executed"]]), float(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    m = re.match(r"\s*(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)", r[ix["Source"]])
    if not m:
        continue
    op = m.group(1) + "".join(x for x in re.findall(r"\.[A-Za-z0-9_]+", m.group(2)) if x in (".128", ".64", ".WIDE", ".x32", ".x8"))
    hist[op] += n; samp[op] += s; total += n; tot_s += s
print(f"total warp-inst {total:.0f} per warp-tile {total / tiles:.1f} samples {tot_s:.0f}")
for op, n in hist.most_common(top):
    print(f"{n / tiles:8.1f} {op:12s} samples {100 * samp[op] / max(tot_s, 1):5.1f}%")

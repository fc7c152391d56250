#!/usr/bin/env python
"""usage: sass_mix.py <cubin|so> <function-substring> [top]  -> SASS instruction histogram of matching kernels"""
import collections
import re
import subprocess
import sys

path, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
on, hist, total, samples = False, collections.Counter(), 0, collections.defaultdict(list)
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
        if on:
            print(line.strip())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)", line)
    if on and m:
        op = m.group(1) + "".join(s for s in re.findall(r"\.[A-Za-z0-9_]+", m.group(2)) if s in (".128", ".64", ".WIDE"))
        hist[op] += 1
        total += 1
        if len(samples[op]) < 2:
            samples[op].append(line.split("*/", 1)[1].split("/*")[0].strip())
print("total", total)
for op, c in hist.most_common(top):
    print(f"{c:6d} {op:14s} {samples[op][0][:90]}")

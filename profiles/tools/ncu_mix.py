#!/usr/bin/env python
"""Executed-instruction mix of one ncu report (SASS opcode histogram weighted by executions).

    python profiles/tools/ncu_mix.py REPORT.ncu-rep
"""
import collections
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
mix = collections.Counter()
samples = collections.Counter()
for r in data:
    src = r[ix["Source"]].split()
    op = src[1] if src[0].startswith("@") else src[0]
    op = op.split(".")[0]
    mix[op] += int(r[ix["Instructions Executed"]])
    samples[op] += int(r[ix["# Samples"]])
tot = sum(mix.values())
ts = sum(samples.values())
print(f"total warp-instructions {tot}, samples {ts}")
for op, n in mix.most_common(40):
    print(f"  {op:12s} {n:10d} {100 * n / tot:5.1f}%   samples {100 * samples[op] / ts:5.1f}%")

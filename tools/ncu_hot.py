#!/usr/bin/env python
"""Hot source lines of one kernel: python tools/ncu_hot.py rep.ncu-rep kernel_regex [topN]"""
import csv, subprocess, sys, collections, re
rep, kr = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kr],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
# the csv has one block per kernel launch: take the first block
rows = list(csv.reader(lines))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] in ("Address", "#")][0]
h = rows[hdr_i]
print(h[:12])

#!/usr/bin/env python
"""Hot CUDA source lines of one kernel (first matching launch) from an .ncu-rep captured with --import-source on:
python tools/ncu_hot.py rep.ncu-rep kernel_regex [topN] [sort: inst|samples]"""
import csv, subprocess, sys
rep, kr = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
key = sys.argv[4] if len(sys.argv) > 4 else "inst"
import os
sel = ["--launch-skip", os.environ["NCU_LAUNCH"], "--launch-count", "1"] if os.environ.get("NCU_LAUNCH") else \
      ["--kernel-name", "regex:" + kr]          # NCU_LAUNCH=<index>: pick one launch by position (templated kernels)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + sel,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
func_pat = sys.argv[5] if len(sys.argv) > 5 else ""     # substring of the "Function Name" row (templated kernels)
out, fname, h, kernels, func_ok = [], "", None, 0, True
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1:
            break
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func_ok = func_pat in r[1]
        continue
    if not func_ok:
        continue
    if r[0] == "Line No":
        h = r
        continue
    if h and r[0].strip().isdigit():
        d = dict(zip(h[4:], r[4:]))
        try:
            inst = int(d["Instructions Executed"]); smp = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        out.append((inst, smp, fname, int(r[0]), r[1].strip()[:110], d))
tot_i = sum(o[0] for o in out) or 1
tot_s = sum(o[1] for o in out) or 1
out.sort(key=lambda o: -(o[0] if key == "inst" else o[1]))
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
for inst, smp, f, ln, src, d in out[:top]:
    st = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s   [%s]" % (100.0 * inst / tot_i, 100.0 * smp / tot_s, f, ln, src,
                                                       ", ".join("%s=%d" % (k[6:], v) for k, v in st)))

#!/usr/bin/env python
"""Per-source-line hot spots of one kernel of an .ncu-rep taken with --import-source on:
python tools/ncu_lines.py rep.ncu-rep <kernel-id (0-based index in the report)> [top N]
Prints, per CUDA source line: warp-stall samples, instructions executed, and the dominant stall reasons."""
import csv, subprocess, sys, collections
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::" + kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname, func, hdr = None, None, None
agg = collections.OrderedDict()
tot_s = tot_i = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        func = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":           # a CUDA source line with aggregated metrics
        try:
            s = int(r[hdr.index("# Samples")]); ins = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        stalls = {}
        for j, h in enumerate(hdr):
            if h.startswith("stall_") and not h.endswith("(Not Issued)"):
                try:
                    v = int(r[j])
                except ValueError:
                    v = 0
                if v:
                    stalls[h[6:]] = v
        key = (fname, int(r[0]))
        e = agg.setdefault(key, [0, 0, r[1].strip(), collections.Counter()])
        e[0] += s; e[1] += ins; e[3].update(stalls)
        tot_s += s; tot_i += ins
print(func, "samples", tot_s, "warp-instructions", tot_i)
for (f, ln), (s, ins, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% ins  %-14s:%-4d %-70s %s" % (100.0 * s / max(1, tot_s), 100.0 * ins / max(1, tot_i), f, ln, src[:70],
                                                      " ".join("%s=%d" % kv for kv in st.most_common(3))))

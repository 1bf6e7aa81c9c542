#!/usr/bin/env python
"""profiles/ncu_traffic.json from an ncu full-set summary CSV (tools/ncu_summary.py):
python tools/ncu_traffic.py profiles/<summary>.csv <frames per launch> "<how it was captured>" """
import csv, json, os, re, sys
src, frames, how = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(open(src)))
h = rows[0]
ki, ti, ri, wi = (h.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"))
units = rows[1]
def conv(v, u):
    v = float(v.replace(",", ""))
    return v * {"Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6, "us": 1.0, "ms": 1e3, "ns": 1e-3}[u]
kern = {}
for r in rows[2:]:
    name = re.sub(r"^void ", "", r[ki]).split("(")[0]
    e = kern.setdefault(name, {"launches": 0, "dram_read_mb": 0.0, "dram_write_mb": 0.0, "time_us": 0.0})
    e["launches"] += 1
    e["dram_read_mb"] += conv(r[ri], units[ri]); e["dram_write_mb"] += conv(r[wi], units[wi]); e["time_us"] += conv(r[ti], units[ti])
for e in kern.values():
    for k in ("dram_read_mb", "dram_write_mb", "time_us"):
        e[k] = round(e[k], 2)
def pick(prefix):
    return next(v for k, v in kern.items() if k.startswith(prefix))
def total(prefixes):
    sel = [v for k, v in kern.items() if any(k.startswith(p) for p in prefixes)]
    return round(sum(v["dram_read_mb"] for v in sel), 2), round(sum(v["dram_write_mb"] for v in sel), 2)
img, ccl, edg = pick("k_decimate_threshold"), pick("k_cc_local"), pick("k_edges")
ccr, ccw = total(["k_cc_local", "k_cc_boundary", "k_cc_sizes", "k_cc_dense"])
out = {"source": "%s (%s)" % (src, how), "frames_per_launch": frames, "kernels": kern,
       "image": {"kernel": "k_decimate_threshold", "dram_read_mb": img["dram_read_mb"], "dram_write_mb": img["dram_write_mb"]},
       "cc": {"kernel": "k_cc_local", "dominant_kernel": {"dram_read_mb": ccl["dram_read_mb"], "dram_write_mb": ccl["dram_write_mb"]},
              "dram_read_mb": ccr, "dram_write_mb": ccw},
       "edges": {"kernel": "k_edges", "dram_read_mb": edg["dram_read_mb"], "dram_write_mb": edg["dram_write_mb"]}}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps({k: out[k] for k in ("image", "cc", "edges")}))

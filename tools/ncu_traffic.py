#!/usr/bin/env python
"""profiles/ncu_traffic.json from an ncu full-set summary CSV (tools/ncu_summary.py):
python tools/ncu_traffic.py profiles/<summary>.csv <frames per launch> <config C1..C5> "<how it was captured>"
The file is stamped with the hash of the kernel sources it was captured from (bench.csrc_sha); bench.py uses the DRAM
bytes only when that hash equals the hash of the sources of the build it is running."""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_sha
src, frames, config, how = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
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
    if not name.startswith("k_fit_quads"):
        name = name.split("<")[0]            # the names agpu_get_kernel_table uses
    e = kern.setdefault(name, {"launches": 0, "dram_read_mb": 0.0, "dram_write_mb": 0.0, "time_us": 0.0})
    e["launches"] += 1
    e["dram_read_mb"] += conv(r[ri], units[ri]); e["dram_write_mb"] += conv(r[wi], units[wi]); e["time_us"] += conv(r[ti], units[ti])
for e in kern.values():
    for k in ("dram_read_mb", "dram_write_mb", "time_us"):
        e[k] = round(e[k], 2)
out = {"source": "%s (%s)" % (src, how), "csrc_sha": csrc_sha(), "config": config, "frames_per_launch": frames, "kernels": kern}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps({k: (v["dram_read_mb"], v["dram_write_mb"], v["time_us"]) for k, v in kern.items()}))

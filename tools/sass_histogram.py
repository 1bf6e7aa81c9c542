#!/usr/bin/env python
"""SASS opcode histogram of the hot kernels of libaprilgpu.so (cuobjdump -sass; no GPU needed):
python tools/sass_histogram.py > profiles/<name>.txt
Puts on record which instructions the kernels are made of: 128-bit loads (LDG.E.128), packed-half min/max/compare
(HMNMX2 / HSET2) in the threshold kernel, DP2A in the BGR front end, shared-memory atomics (ATOMS), FP64 (DADD / DMUL /
DFMA), warp primitives (SHFL / VOTE / MATCH) -- and that there are no tensor-core (HMMA / UTCMMA) or TMA (UTMALDG)
instructions: nothing here is a dense contraction or a bulk tile copy."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "aprilslam_b200", "libaprilgpu.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1)).split("(")[0].replace("void ", "")
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        kern[cur][m.group(1)] += 1
WATCH = ["LDG.E.128", "LDG.E.64", "LDG.E.U8", "STG.E.128", "STG.E.64", "HMNMX2", "HSET2", "IDP.2A", "IDP.4A", "PRMT", "SHF", "REDUX", "ATOMS", "ATOMG", "RED",
         "SHFL", "VOTE", "MATCH", "DADD", "DMUL", "DFMA", "MUFU", "BAR", "HMMA", "UTCMMA", "UTMALDG", "LDL", "STL"]
hot = [k for k in kern if any(k.startswith(p) for p in ("k_decimate_threshold<1, 4, false, 1>", "k_decimate_threshold<1, 3, false, 3>",
       "k_decimate_blur_strip<1, 3, false>", "k_decimate_blur<1>", "k_cc_local<false>", "k_cc_boundary<8>", "k_edges<2>", "k_sort_scatter",
       "k_fit_quads<2>", "k_decode_quads", "k_reconcile", "k_pose"))]
print("# cuobjdump -sass aprilslam_b200/libaprilgpu.so (sm_100a), opcode counts per kernel (static instruction count)")
for k in hot:
    c = kern[k]
    tot = sum(c.values())
    grp = collections.Counter()
    for op, n in c.items():
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("LDG.E.128", "LDG.E.64", "LDG.E.U8", "STG.E.128", "STG.E.64") and op.startswith(w.split(".E")[0]) and w.split("E.")[1] in op.split(".")):
                grp[w] += n
                break
    print("\n%s: %d instructions" % (k, tot))
    print("   watched : " + "  ".join("%s=%d" % (w, grp[w]) for w in WATCH if grp[w]))
    print("   top     : " + "  ".join("%s=%d" % kv for kv in c.most_common(14)))
allops = collections.Counter()
for c in kern.values():
    allops.update(c)
tc = [op for op in allops if op.startswith(("HMMA", "IMMA", "UTCMMA", "UTCHMMA", "UTMALDG", "UTMASTG", "TCGEN", "QGMMA"))]
print("\n# tensor-core / TMA opcodes anywhere in the library: %s" % (tc or "none"))

#!/usr/bin/env python
"""Top stall sites of an `ncu --set full --import-source on` capture (source page), run here without a GPU:
   python tools/ncu_stalls.py gpurun_out/prof_x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[isamp]), r[ia], r[isrc], r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print(f"# {rows[0][1][:100] if rows and len(rows[0]) > 1 else ''}\n# total samples {tot}, {len(data)} instructions")
agg = {}
for n, a, src, r in data:
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            agg[hdr[i]] = agg.get(hdr[i], 0) + v
print("# stall mix:", ", ".join(f"{k}={100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, a, src, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{n:6d} {100 * n / tot:5.1f}% {a[-5:]} {src[:80]:80s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")

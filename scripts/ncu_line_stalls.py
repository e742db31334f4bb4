#!/usr/bin/env python
"""Per source line of one ncu --set full --import-source capture: warp-state samples split by stall reason, instructions executed,
local-memory traffic. Usage: python scripts/ncu_line_stalls.py <report.ncu-rep> [top-N]"""
import csv
import io
import os
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, hdr, agg = None, None, {}
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[0] == "Function Name" or len(r) < 8 or r[2] != "-":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    agg[(cur, ln)] = (dict(zip(hdr[4:], r[4:])), r[1].strip()[:90])
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(d.get("# Samples") or 0) for d, _ in agg.values()) or 1
glob = {k: sum(int(d.get(k) or 0) for d, _ in agg.values()) for k in reasons}
print("all lines:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(glob.items(), key=lambda kv: -kv[1]) if v * 100 > tot))
print("instructions executed:", sum(int(d.get("Instructions Executed") or 0) for d, _ in agg.values()), " local sectors:", sum(int(d.get("L2 Theoretical Sectors Local") or 0) for d, _ in agg.values()))
for (f, ln), (d, text) in sorted(agg.items(), key=lambda kv: -int(kv[1][0].get("# Samples") or 0))[:topn]:
    s = int(d.get("# Samples") or 0)
    top = sorted(((int(d.get(k) or 0), k[6:]) for k in reasons), reverse=True)[:3]
    print(f"{100 * s / tot:5.1f}%  inst {int(d.get('Instructions Executed') or 0):>9d}  {f}:{ln:<4d} " + " ".join(f"{n}:{100 * v / max(s, 1):.0f}%" for v, n in top if v) + f"  | {text}")

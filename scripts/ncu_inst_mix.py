#!/usr/bin/env python
"""Instruction mix of one ncu capture: warp-level instructions executed per source line and per SASS opcode, with stall samples.
Usage: python scripts/ncu_inst_mix.py <file.ncu-rep> [top]"""
import csv, io, os, subprocess, sys
from collections import defaultdict

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    cur = None; hdr = None
    per_line = defaultdict(lambda: [0, 0, ""]); per_op = defaultdict(lambda: [0, 0])
    seen_addr = set()
    curline = None
    for r in csv.reader(io.StringIO(out)):
        if not r: continue
        if r[0] == "File Path": cur = os.path.basename(r[1]); continue
        if r[0] == "Line No": hdr = r; continue
        if hdr is None or r[0] == "Function Name" or len(r) < 8: continue
        dd = dict(zip(hdr[4:], r[4:]))
        if r[2] == "-":
            try: curline = (cur, int(r[0]))
            except ValueError: curline = None
            if curline: per_line[curline][2] = r[1].strip()[:110]
            continue
        if r[2] in ("...", ""): continue
        try: inst = int(dd.get("Instructions Executed") or 0); samp = int(dd.get("# Samples") or 0)
        except ValueError: continue
        if curline:
            per_line[curline][0] += inst; per_line[curline][1] += samp
        if r[2] in seen_addr: continue  # an instruction inlined from several lines is listed once per file section
        seen_addr.add(r[2])
        op = r[3].strip().split()
        if op and op[0].startswith("@"): op = op[1:]
        name = op[0].split(".")[0] if op else "?"
        per_op[name][0] += inst; per_op[name][1] += samp
    tot = sum(v[0] for v in per_op.values()) or 1; tots = sum(v[1] for v in per_op.values()) or 1
    print(f"{rep}: {tot} warp instructions executed, {tots} samples")
    print("-- by opcode"); 
    for k, v in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {100*v[0]/tot:5.1f}% inst  {100*v[1]/tots:5.1f}% samples  {v[0]:>11d}  {k}")
    tl = sum(v[0] for v in per_line.values()) or 1
    print("-- by source line (instructions)")
    for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {100*v[0]/tl:5.1f}% inst  {100*v[1]/tots:5.1f}% samples  {k[0]}:{k[1]}  {v[2]}")

main()

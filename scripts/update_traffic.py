#!/usr/bin/env python
"""profiles/ncu_traffic.json entries for cfg2 / cfg4 from the captures of scripts/profile_other.sh: DRAM bytes (read + write) per launch and
the kernel's duration under ncu. Usage: python scripts/update_traffic.py <tag>"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
KEYS = {"k1": "fk_residual_jacobian", "gram": "jtj_jtr", "chol": "cholesky_update", "k2": "gram_cholesky"}
BATCH = {"cfg2": 4096, "cfg4": 2048}


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.load(open(tj))
lines = []
for wl in ("cfg2", "cfg4"):
    entry = {}
    for short, name in KEYS.items():
        rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{wl}_{short}.ncu-rep")
        if not os.path.exists(rep):
            continue
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h, u, v = rows[0], rows[1], rows[2]
        m = {n: (v[i], u[i]) for i, n in enumerate(h)}
        entry[name] = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
        lines.append(f"{wl} {name:22s} {m['Kernel Name'][0][:60]:60s} duration {m['gpu__time_duration.sum'][0]} {m['gpu__time_duration.sum'][1]}  DRAM {entry[name] / 1e6:.1f} MB  "
                     f"issue-active {float(m['smsp__issue_active.avg.pct_of_peak_sustained_active'][0]):.1f} %  registers {m['launch__registers_per_thread'][0]}")
    if entry:
        d[f"{wl}:{BATCH[wl]}"] = {"dram_bytes_per_launch": entry, "source": f"profiles/{tag}_other_workloads.txt"}
json.dump(d, open(tj, "w"), indent=1)
open(os.path.join(ROOT, "profiles", f"{tag}_other_workloads.txt"), "w").write(
    f"ncu --metrics dram__bytes_read/write.sum, gpu__time_duration.sum, ... --clock-control none, one launch per kernel (scripts/profile_other.sh {tag}); cfg2 = 4096 x humanoid72 (24 Position), cfg4 = 2048 x bodyhands300\n" + "\n".join(lines) + "\n")
print("\n".join(lines))

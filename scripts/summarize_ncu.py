#!/usr/bin/env python
"""Turn the ncu captures of scripts/profile_gpu.sh (gpurun_out/<tag>_*.ncu-rep, <tag>_launches.csv) into the committed
text summaries under profiles/: key metrics per kernel, the hottest source lines, the launch list with each kernel's share.
Usage: python scripts/summarize_ncu.py <tag> [<out-prefix>]"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def kernel_summary(rep, lines):
    rows = ncu_csv(rep, "raw")
    if len(rows) < 3:
        lines.append(f"  (no data in {rep})")
        return None
    h, u, v = rows[0], rows[1], rows[2]
    d = {n: (v[i], u[i]) for i, n in enumerate(h)}
    lines.append(f"kernel: {d.get('Kernel Name', ('?', ''))[0]}   grid {d.get('launch__grid_size', ('?',''))[0]} x block {d.get('launch__block_size', ('?',''))[0]}")
    for k in KEYS:
        if k in d:
            lines.append(f"  {k:80s} {d[k][0]:>16s} {d[k][1]}")
    for n in h:
        if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio") and "not_issued" not in n:
            try:
                if float(d[n][0]) >= 0.3:
                    lines.append(f"  stall {n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {float(d[n][0]):8.2f} warps per issue")
            except ValueError:
                pass
    traffic = None
    if "dram__bytes_read.sum" in d:
        traffic = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
        lines.append(f"  => DRAM traffic per launch {traffic / 1e6:.1f} MB")
    # hottest source lines
    src = ncu_csv(rep, "source", ("--print-source", "sass,cuda"))
    cur, hdr, agg = None, None, {}
    for r in src:
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
        dd = dict(zip(hdr[4:], r[4:]))
        agg[(cur, ln)] = (int(dd.get("# Samples") or 0), int(dd.get("Instructions Executed") or 0), r[1].strip()[:100])
    tot = sum(x[0] for x in agg.values()) or 1
    lines.append("  hottest source lines (warp-state samples):")
    for (f, ln), x in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]:
        lines.append(f"    {100.0 * x[0] / tot:5.1f}%  inst {x[1]:>10d}  {f}:{ln}  {x[2]}")
    return traffic


def launches_summary(path, lines):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    per = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[4].split("(")[0]
        try:
            ns = float(r[-1].replace(",", ""))
        except ValueError:
            continue
        unit = r[-2]
        ns *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)  # gpu__time_duration unit
        per[name][0] += 1
        per[name][1] += ns
    tot = sum(v[1] for v in per.values()) or 1
    lines.append(f"launch list: {len(rows)} launches, {tot / 1e6:.3f} ms total (under ncu: cold-cache, serialised; shares are what matter)")
    for name, (cnt, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"  {100 * ns / tot:5.1f}%  {cnt:4d} x {ns / cnt / 1e3:10.1f} us  {name}")


def main():
    tag = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else tag
    g = os.path.join(ROOT, "gpurun_out")
    lines = [f"ncu summary '{tag}' (scripts/profile_gpu.sh: bench.py --steps 1 --warmup 1, default workload cfg3-shard, 8192 instances, 1 B200)", ""]
    lp = os.path.join(g, f"{tag}_launches.csv")
    if os.path.exists(lp):
        launches_summary(lp, lines)
        lines.append("")
    traffic = {}
    # r02 default path: k1 = sweepKernel, k2 = gramCholeskyKernel; the other captures (three-kernel path, persistent kernel) are for comparison
    for key, name in (("k1", "fk_residual_jacobian"), ("k2", "gram_cholesky"), ("gram", "jtj_jtr"), ("chol", "cholesky_update"), ("persistent", "persistent_solve"), ("k3", "cholesky_update")):
        rep = os.path.join(g, f"{tag}_{key}.ncu-rep")
        if os.path.exists(rep):
            lines.append(f"== {key} ({name}) -- ncu --set full --clock-control none, one launch")
            t = kernel_summary(rep, lines)
            if t is not None:
                traffic[name] = t
            lines.append("")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{out}_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
    tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    d = json.load(open(tj)) if os.path.exists(tj) else {}
    d["cfg3-shard:8192"] = {"dram_bytes_per_launch": traffic, "source": f"profiles/{out}_ncu_summary.txt"}
    json.dump(d, open(tj, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()

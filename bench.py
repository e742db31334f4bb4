#!/usr/bin/env python
"""bench.py — GN iterations/sec of the batched 72-joint IK hot path (BASELINE.json metric).

A *step* is one batched solve: every instance of the per-GPU shard runs ``ITERS`` damped Gauss-Newton
iterations (minIterations = maxIterations = ITERS, so no instance stops early) of
FK sweep -> residual/Jacobian -> JtJ/Jtr -> damped Cholesky -> update.

Default workload "cfg3-shard": the per-GPU shard of BASELINE.json configs[2] — the configuration the
metric/target is quoted on ("72-joint / 128-residual batch at 8xB200") — 8192 x humanoid72 with 24
Position + 6 Orientation constraints (m = 126 -> 128 rows, n = 220), constant damping 0.05, weak
scaling (8 GPUs = 65536 instances = cfg3 exactly). ``--workload cfg2`` / ``cfg4`` select the other
single-GPU configs.

    value   : whole-job GN iterations/s with parameters and targets resident in HBM
    e2e     : same metric through mb2_solver_solve with HOST (pinned) buffers: H2D of targets and initial
              parameters and D2H of solved parameters + per-instance results inside the timed region
    roofline: the JtJ kernel (the kernel BASELINE's metric names), algorithmic m*n*(n+1) FLOP per
              instance-iteration over its CUDA-event time, against the measured tensor peak
    cpu_baseline / --impl reference: the oracle restatement of the reference's CPU solver (one solver per
              instance over all host threads, as tensor_ik.cpp:127), bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITERS = 10
WORKLOADS = {
    "cfg3-shard": dict(batch=8192, rig="humanoid72", orientation=True, desc="8192/GPU x humanoid72, 24 Position + 6 Orientation (m=126->128, n=220), lambda=0.05"),
    "cfg2": dict(batch=4096, rig="humanoid72", orientation=False, desc="4096/GPU x humanoid72, 24 Position (m=72, n=220), lambda=0.05"),
    "cfg4": dict(batch=2048, rig="bodyhands300", orientation=False, desc="2048/GPU x bodyhands300, 200 Position (m=600, n=424), lambda=0.05"),
}


def make_problem(workload, batch, seed_offset=0):
    from momentum_b200.problems import bodyhands_problem, humanoid_problem

    w = WORKLOADS[workload]
    if w["rig"] == "humanoid72":
        return humanoid_problem(batch, seed=12347 + seed_offset, orientation=w["orientation"])
    return bodyhands_problem(batch, seed=12349 + seed_offset)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port, float) on the host cores."""
    if rank != 0:
        return
    from oracle.binding import OracleFunction, hardware_threads

    threads = hardware_threads()
    sample = max(threads * 4, 64)
    ch, efs, theta0, _ = make_problem(args.workload, sample)
    orc = OracleFunction(ch, efs, "float32")
    kw = dict(threads=threads, min_iterations=ITERS, max_iterations=ITERS, threshold=1.0, regularization=0.05, final_errors=False)
    for _ in range(args.warmup):
        orc.solve_batch(theta0, **kw)
    t0 = time.perf_counter()
    its = 0
    for _ in range(args.steps):
        r = orc.solve_batch(theta0, **kw)
        its += int(r["iterations"].sum())
    dt = time.perf_counter() - t0
    value = its / dt
    line = {"impl": "reference", "metric": "GN iterations/sec (batched 72-joint IK)", "value": value, "unit": "GN it/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": args.workload, "desc": WORKLOADS[args.workload]["desc"], "iterations_per_solve": ITERS},
            "cpu_baseline": {"value": value, "unit": "GN it/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} instances x {ITERS} GN iterations per step (oracle restatement, float, one solver per instance over {threads} threads)"},
            "e2e": {"value": value, "unit": "GN it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg3-shard", choices=sorted(WORKLOADS))
    ap.add_argument("--batch-per-gpu", type=int, default=0)
    ap.add_argument("--jtj-mode", type=int, default=0)
    ap.add_argument("--cholesky-mode", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from momentum_b200 import solver as ms

    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.batch_per_gpu or WORKLOADS[args.workload]["batch"]
    ch, efs, theta0, _ = make_problem(args.workload, B, seed_offset=1000 * rank)  # each rank owns a different shard
    n = ch.num_params
    fn = ms.SkeletonSolverFunction(ch, B, efs, device=local_rank)
    fn.upload_targets()
    opts = ms.GaussNewtonSolverOptions(min_iterations=ITERS, max_iterations=ITERS, threshold=1.0, regularization=0.05, jtj_mode=args.jtj_mode, cholesky_mode=args.cholesky_mode)
    solver = ms.GaussNewtonSolver(opts, fn)
    m_rows = sum(3 * len(e.parents) if e.kind == 0 else 9 * len(e.parents) for e in efs)

    work_stream = torch.cuda.Stream()  # a real (non-NULL) stream: NULL means "the handle's own stream" in the C-ABI
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    theta0_dev = torch.from_numpy(theta0.astype(np.float32)).cuda()
    theta_dev = torch.empty_like(theta0_dev)
    # pinned host buffers for the e2e leg
    theta0_pin = torch.from_numpy(theta0.astype(np.float32)).pin_memory()
    theta_pin = torch.empty_like(theta0_pin).pin_memory()
    target_pins = [torch.from_numpy(np.ascontiguousarray(e.targets, np.float32)).pin_memory() for e in efs]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        theta_dev.copy_(theta0_dev)
        solver.solve_device(theta_dev.data_ptr(), stream)

    def e2e_step():
        for idx, tp in enumerate(target_pins):
            fn._check(fn._L.mb2_set_targets(fn._h, idx, ms.C.cast(tp.data_ptr(), ms._fp)))
        theta_pin.copy_(theta0_pin)
        solver.solve_host_pointer(theta_pin.data_ptr())
        return solver.get_results()

    # ---- device-resident timing (value) ----
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()  # flush L2 between timed iterations
        torch.cuda.synchronize()
        ev0.record()
        device_step()
        ev1.record()
        torch.cuda.synchronize()
        total_ms += ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    barrier()
    res = solver.get_results()
    its_per_step = int(res["iterations"].sum())
    total_iter, launches = solver.get_counters()
    # the one collective of the path: aggregate iterations / residual norm (SUM) and elapsed device time (MAX over ranks)
    from momentum_b200.distributed import aggregate_solve_stats

    its_total, err_total, max_ms = aggregate_solve_stats(float(its_per_step), float(res["errors"].sum()), total_ms, device="cuda")
    value = its_total * args.steps / (max_ms * 1e-3)

    # ---- per-kernel times for the roofline (profiling mode: events around every launch) ----
    solver.set_profiling(True)
    device_step()
    torch.cuda.synchronize()
    solver.get_results()
    phase_ms, phase_launches = solver.get_phase_times()
    solver.set_profiling(False)

    # ---- e2e through the host-buffer C-ABI call ----
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = its_total * args.steps / te.item()
    h2d = int(theta0_pin.numel() * 4 + sum(tp.numel() * 4 for tp in target_pins))
    d2h = int(theta_pin.numel() * 4 + B * (8 + 4 + 4))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    tf32_peak = 0.5 * peaks["bf16"]  # TF32 dense = half the bf16 rate (B200_PROFILING.md table); bf16 figure is the measured cuBLAS burst
    jtj_flops = float(m_rows) * n * (n + 1)
    jtj_ms = phase_ms[1] / max(1, phase_launches[1])
    achieved = jtj_flops * B / (jtj_ms * 1e-3) / 1e12 if jtj_ms > 0 else 0.0
    sweep_ms = phase_ms[0] / max(1, phase_launches[0])
    chol_ms = phase_ms[2] / max(1, phase_launches[2])
    line = {
        "metric": "GN iterations/sec (batched 72-joint IK)", "value": value, "unit": "GN it/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": WORKLOADS[args.workload]["desc"], "batch_per_gpu": B, "global_batch": B * world, "iterations_per_solve": ITERS,
                   "rows_m": m_rows, "params_n": n, "parallelism": f"dp{world} (instances sharded, no data-path collective)",
                   "jtj_mode": args.jtj_mode, "cholesky_mode": args.cholesky_mode, "l2": "256 MB buffer written between timed steps (L2 flush)"},
        "solves_per_sec": value / ITERS, "aggregate_final_error": err_total,
        "e2e": {"value": e2e_value, "unit": "GN it/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "JtJ/Jtr", "bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                     "peak_source": f"0.5 x {peaks['src']} bf16 cuBLAS burst ({peaks['bf16']} TF/s) = TF32 dense", "traffic": None,
                     "ms_per_launch": jtj_ms, "algorithmic_flops_per_instance": jtj_flops},
        "kernels_ms_per_iteration": {"fk_residual_jacobian": sweep_ms, "jtj_jtr": jtj_ms, "cholesky_update": chol_ms},
    }
    if not args.no_cpu_baseline:
        from oracle.binding import OracleFunction, hardware_threads

        threads = hardware_threads()
        sample = max(threads * 4, 64)
        ch_s, efs_s, th_s, _ = make_problem(args.workload, sample)
        orc = OracleFunction(ch_s, efs_s, "float32")
        kw = dict(threads=threads, min_iterations=ITERS, max_iterations=ITERS, threshold=1.0, regularization=0.05, final_errors=False)
        orc.solve_batch(th_s, **kw)
        reps, t0 = 0, time.perf_counter()
        its = 0
        while time.perf_counter() - t0 < 10.0:
            its += int(orc.solve_batch(th_s, **kw)["iterations"].sum())
            reps += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": its / dt, "unit": "GN it/s", "cores": threads, "kind": "port",
                                "sample": f"{reps} x {sample} instances x {ITERS} GN iterations (oracle restatement of the reference solver, float, one solver per instance over {threads} threads)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITERS = 10
WORKLOADS = {
    "cfg3-shard": dict(batch=8192, rig="humanoid72", orientation=True, rows_m=126, params_n=220, desc="8192/GPU x humanoid72, 24 Position + 6 Orientation (m=126->128, n=220), lambda=0.05"),
    "cfg2": dict(batch=4096, rig="humanoid72", orientation=False, rows_m=72, params_n=220, desc="4096/GPU x humanoid72, 24 Position (m=72, n=220), lambda=0.05"),
    "cfg4": dict(batch=2048, rig="bodyhands300", orientation=False, rows_m=600, params_n=424, desc="2048/GPU x bodyhands300, 200 Position (m=600, n=424), lambda=0.05"),
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


def measured_traffic(workload, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures (profiles/ncu_traffic.json);
    only valid for the exact workload / batch they were captured on."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return {}
    d = json.load(open(p))
    return d.get(f"{workload}:{batch}", {}).get("dram_bytes_per_launch", {})


class ClockSampler:
    """SM clock / clock-event reasons sampled DURING the timed regions (B200_PROFILING.md recipe). NVML is polled in-process
    every 5 ms (the timed region of the default run is ~0.1 s, shorter than one nvidia-smi start-up)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index, self.sm, self.mask, self.max_mhz, self.handle, self.nvml = index, [], 0, None, None, None
        self.stop_flag = threading.Event()
        self.thread = None
        self.error = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception as e:  # noqa: BLE001
            self.error = f"nvml unavailable: {e}"

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception as e:  # noqa: BLE001
                self.error = str(e)
                return
            time.sleep(0.005)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.error or "nvml unavailable"], "samples": 0}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.sm)}


def bench_config(args, world):
    """The `config` dict of both arms' JSON lines (the driver compares them key by key): the workload as named in BASELINE.json, never
    anything measured. The CPU arm times a bounded sample of this workload (its size is in cpu_baseline.sample)."""
    w = WORKLOADS[args.workload]
    B = args.batch_per_gpu or w["batch"]
    return {"workload": args.workload, "desc": w["desc"], "batch_per_gpu": B, "global_batch": B * world, "iterations_per_solve": ITERS,
            "rows_m": w["rows_m"], "params_n": w["params_n"], "parallelism": f"dp{world} (instances sharded, no data-path collective)",
            "jtj_mode": args.jtj_mode, "cholesky_mode": args.cholesky_mode, "l2": "256 MB buffer written between timed steps (L2 flush)"}


def time_cpu_arm(workload, seconds, threads=None):
    """The reference's CPU algorithm (oracle restatement, float, -march=native timing build, one solver per instance over the host
    threads this process may use — affinity mask capped by the cgroup quota — as tensor_ik.cpp:127 does with dispenso) on a bounded
    sample of the workload. Returns (it/s, threads, sample description, single-thread it/s)."""
    from oracle.binding import OracleFunction, hardware_threads

    threads = threads or hardware_threads()
    sample = max(threads * 4, 64)
    ch, efs, theta0, _ = make_problem(workload, sample)
    orc = OracleFunction(ch, efs, "float32", native=True)
    kw = dict(min_iterations=ITERS, max_iterations=ITERS, threshold=1.0, regularization=0.05, final_errors=False)
    orc.solve_batch(theta0, threads=threads, **kw)
    reps, its, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        its += int(orc.solve_batch(theta0, threads=threads, **kw)["iterations"].sum())
        reps += 1
    dt = time.perf_counter() - t0
    # single-thread figure on a few instances
    n1 = min(sample, 8)
    t1 = time.perf_counter()
    its1 = int(orc.solve_batch(theta0[:n1], threads=1, **kw)["iterations"].sum()) if n1 == sample else None
    if its1 is None:
        ch1, efs1, th1, _ = make_problem(workload, n1)
        o1 = OracleFunction(ch1, efs1, "float32", native=True)
        t1 = time.perf_counter()
        its1 = int(o1.solve_batch(th1, threads=1, **kw)["iterations"].sum())
    single = its1 / (time.perf_counter() - t1)
    build = "-O3 -march=native, FMA + reassociation, row-major blocked LLT" if getattr(orc._L, "is_native", False) else "portable x86-64-v3 checker build"
    desc = (f"{reps} x {sample} instances x {ITERS} GN iterations (oracle restatement of the reference solver, float, {build}; one solver per "
            f"instance over {threads} host threads; single thread: {single:.0f} GN it/s)")
    return its / dt, threads, desc, single, sample


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port, float) on the host cores."""
    if rank != 0:
        return
    budget = max(5.0, min(20.0, 2.0 * (args.steps + args.warmup)))  # bounded: the whole run ends within a few minutes whatever K / W are
    t0 = time.perf_counter()
    value, threads, desc, single, sample = time_cpu_arm(args.workload, budget)
    dt = time.perf_counter() - t0
    line = {"impl": "reference", "metric": "GN iterations/sec (batched 72-joint IK)", "value": value, "unit": "GN it/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sample * ITERS / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(args, world),
            "cpu_baseline": {"value": value, "unit": "GN it/s", "cores": threads, "kind": "port", "sample": desc, "single_thread_value": single},
            "e2e": {"value": value, "unit": "GN it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "wall_s": dt}
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
    assert m_rows == WORKLOADS[args.workload]["rows_m"] and n == WORKLOADS[args.workload]["params_n"]

    work_stream = torch.cuda.Stream()  # a real (non-NULL) stream: NULL means "the handle's own stream" in the C-ABI
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    theta0_dev = torch.from_numpy(theta0.astype(np.float32)).cuda()
    theta_dev = torch.empty_like(theta0_dev)
    # pinned host buffers for the e2e leg
    theta0_pin = torch.from_numpy(theta0.astype(np.float32)).pin_memory()
    # the solve is in place (like the reference's solve(params)): one pinned in/out buffer per e2e step, filled before the timed region
    n_e2e = max(1, args.warmup // 2) + args.steps
    theta_pins = [theta0_pin.clone().pin_memory() for _ in range(n_e2e)]
    theta_pin = theta_pins[0]
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

    e2e_count = [0]

    def e2e_step():
        buf = theta_pins[e2e_count[0] % n_e2e]
        e2e_count[0] += 1
        for idx, tp in enumerate(target_pins):
            fn._check(fn._L.mb2_set_targets(fn._h, idx, ms.C.cast(tp.data_ptr(), ms._fp)))
        solver.solve_host_pointer(buf.data_ptr())
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
    clocks = sampler.stop()  # sampled across the device-resident and the end-to-end timed regions
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
    hbm_peak = peaks["hbm_gbs"]
    st = solver.get_plan_stats()
    sweep_ms = phase_ms[0] / max(1, phase_launches[0])
    jtj_ms = phase_ms[1] / max(1, phase_launches[1])
    chol_ms = phase_ms[2] / max(1, phase_launches[2])
    target_floats = sum(tp.numel() for tp in target_pins) // B
    # ALGORITHMIC bytes / flops per instance and launch (DESIGN.md section 4 derives each figure)
    k1_bytes = 4.0 * (n + target_floats + st["jacobian_nonzeros"] + m_rows) + 8.0
    jtj_flops = float(m_rows) * n * (n + 1)
    k2_bytes = 4.0 * (m_rows * (st["jacobian_columns"] + 1) + (st["normal_parameters"] + 1) * (st["normal_parameters"] + 2) / 2)
    k3_entries = st["cholesky_tiles"] * 256 if st["cholesky_tiles"] else st["normal_parameters"] * (st["normal_parameters"] + 1) / 2
    k3_bytes = 4.0 * (k3_entries + 4 * st["normal_parameters"])
    traffic = measured_traffic(args.workload, B)

    def hbm_entry(name, key, bytes_per_instance, ms):
        ach = bytes_per_instance * B / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "peak_source": f"{peaks['src']} copy bandwidth", "traffic": traffic.get(key), "ms_per_launch": ms,
                "algorithmic_bytes_per_instance": bytes_per_instance}

    gram = st["strip_floats"] > 0  # tile-sparse Gram path: strips in, tiles out, no dense J / H
    if gram:
        k1_bytes = 4.0 * (n + target_floats + st["jacobian_nonzeros"] + m_rows) + 8.0
        gram_bytes = 4.0 * (st["strip_floats"] + st["cholesky_tiles"] * 256 + 16 * ((st["normal_parameters"] + 15) // 16))
        gram_flops = 2.0 * st["gram_macs"]
        gram_ach = gram_flops * B / (jtj_ms * 1e-3) / 1e12 if jtj_ms > 0 else 0.0
        k2_entry = hbm_entry("gramTilesKernel (tile-sparse J^T J / J^T r, mma.sync 3xTF32 over non-zero strips)", "jtj_jtr", gram_bytes, jtj_ms)
        # the JtJ kernel is credited m n (n + 1) flops per instance whatever it executes; both forms are reported
        dense_eq = jtj_flops * B / (jtj_ms * 1e-3) / 1e12 if jtj_ms > 0 else 0.0
        step_ms = sweep_ms + jtj_ms + chol_ms
        k2_entry.update({"algorithmic_flops_per_instance": gram_flops, "tflops": gram_ach, "dense_equivalent_flops_per_instance": jtj_flops,
                         "dense_equivalent_tflops": dense_eq, "dense_equivalent_frac_of_tf32_peak": dense_eq / tf32_peak,
                         "dense_equivalent_tflops_end_to_end": jtj_flops * B / (step_ms * 1e-3) / 1e12 if step_ms > 0 else 0.0})
    else:
        jtj_ach = jtj_flops * B / (jtj_ms * 1e-3) / 1e12 if jtj_ms > 0 else 0.0
        k2_entry = {"kernel": "jtjTensorKernel (JtJ/Jtr, tcgen05 3xTF32)" if st["jacobian_columns"] + 1 <= 256 and args.jtj_mode in (0, 2, 3) else "jtjSimtKernel",
                    "bound": "tensor", "achieved": jtj_ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": jtj_ach / tf32_peak,
                    "peak_source": f"0.5 x {peaks['src']} bf16 cuBLAS burst ({peaks['bf16']} TF/s) = TF32 dense", "traffic": traffic.get("jtj_jtr"),
                    "ms_per_launch": jtj_ms, "algorithmic_flops_per_instance": jtj_flops, "algorithmic_bytes_per_instance": k2_bytes,
                    "hbm_frac": (k2_bytes * B / (jtj_ms * 1e-3) / 1e9 / hbm_peak) if jtj_ms > 0 else 0.0}
    kernels = [
        hbm_entry("sweepKernel<true> (FK + residual + Jacobian)", "fk_residual_jacobian", k1_bytes, sweep_ms),
        k2_entry,
        hbm_entry("choleskyScheduledKernel (damped LLT + solves + update)" if st["cholesky_tiles"] else "choleskyKernel (dense LLT + solves + update)",
                  "cholesky_update", k3_bytes, chol_ms),
    ]
    dominant = max(kernels, key=lambda k: k["ms_per_launch"])
    line = {
        "metric": "GN iterations/sec (batched 72-joint IK)", "value": value, "unit": "GN it/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, world),
        "solves_per_sec": value / ITERS, "aggregate_final_error": err_total,
        "e2e": {"value": e2e_value, "unit": "GN it/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": dominant,
        "roofline_all_kernels": kernels,
        "kernels_ms_per_iteration": {"fk_residual_jacobian": sweep_ms, "jtj_jtr": jtj_ms, "cholesky_update": chol_ms},
    }
    if not args.no_cpu_baseline:
        value_cpu, threads, desc, single, _ = time_cpu_arm(args.workload, 10.0)
        line["cpu_baseline"] = {"value": value_cpu, "unit": "GN it/s", "cores": threads, "kind": "port", "sample": desc, "single_thread_value": single}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

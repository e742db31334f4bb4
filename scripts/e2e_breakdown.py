#!/usr/bin/env python
"""Runs on the GPU box: where the end-to-end step (host buffers in and out through the C-ABI) spends its time next to the
device-resident solve, call by call (host wall clock with a device synchronize after every call)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from bench import ITERS, make_problem  # noqa: E402
from momentum_b200 import solver as ms  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ch, efs, theta0, _ = make_problem("cfg3-shard", B)
fn = ms.SkeletonSolverFunction(ch, B, efs, device=0)
fn.upload_targets()
solver = ms.GaussNewtonSolver(ms.GaussNewtonSolverOptions(min_iterations=ITERS, max_iterations=ITERS, threshold=1.0, regularization=0.05), fn)
theta_pin = torch.from_numpy(theta0.astype(np.float32)).pin_memory()
bufs = [theta_pin.clone().pin_memory() for _ in range(4)]
target_pins = [torch.from_numpy(np.ascontiguousarray(e.targets, np.float32)).pin_memory() for e in efs]
dev = torch.from_numpy(theta0.astype(np.float32)).cuda()
work = torch.empty_like(dev)
stream = torch.cuda.current_stream().cuda_stream


def timed(label, f, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = f()
    torch.cuda.synchronize()
    acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0
    return r


for rep in range(3):
    acc = {}
    n = 5
    for i in range(n):
        for idx, tp in enumerate(target_pins):
            timed(f"set_targets[{idx}] ({tp.numel() * 4 / 1e6:.2f} MB)", lambda: fn._check(fn._L.mb2_set_targets(fn._h, idx, ms.C.cast(tp.data_ptr(), ms._fp))), acc)
        timed("solve_host (H2D theta + solve + D2H theta + wait)", lambda: solver.solve_host_pointer(bufs[i % 4].data_ptr()), acc)
        timed("get_results", solver.get_results, acc)
        work.copy_(dev)
        timed("solve_device only", lambda: solver.solve_device(work.data_ptr(), stream), acc)
        timed("H2D theta alone (pinned, torch)", lambda: work.copy_(theta_pin, non_blocking=True), acc)
        timed("D2H theta alone (pinned, torch)", lambda: bufs[0].copy_(work, non_blocking=True), acc)
    if rep == 2:
        for k, v in acc.items():
            print(f"{k:60s} {1e3 * v / n:8.3f} ms")
# the whole step as bench.py times it
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(5):
    for idx, tp in enumerate(target_pins):
        fn._check(fn._L.mb2_set_targets(fn._h, idx, ms.C.cast(tp.data_ptr(), ms._fp)))
    solver.solve_host_pointer(bufs[i % 4].data_ptr())
    solver.get_results()
torch.cuda.synchronize()
print(f"{'e2e step as bench.py runs it':60s} {1e3 * (time.perf_counter() - t0) / 5:8.3f} ms")

# kernels timed with CUDA events INSIDE the end-to-end call (profiling level 1): is the drift of the call on the device or on the host?
solver.set_profiling(1)
for i in range(8):
    t0 = time.perf_counter()
    solver.solve_host_pointer(bufs[i % 4].data_ptr())
    wall = 1e3 * (time.perf_counter() - t0)
    solver.get_results()
    ms_, ln = solver.get_phase_times()
    print(f"profiled e2e call {i}: wall {wall:.3f} ms, kernels (events) sweep {ms_[0]:.3f} + gram/chol {ms_[2]:.3f} = {ms_[0] + ms_[1] + ms_[2]:.3f} ms")
solver.set_profiling(0)
for i in range(12):
    t0 = time.perf_counter()
    solver.solve_host_pointer(bufs[i % 4].data_ptr())
    print(f"plain e2e call {i}: wall {1e3 * (time.perf_counter() - t0):.3f} ms")

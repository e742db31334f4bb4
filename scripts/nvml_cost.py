#!/usr/bin/env python
"""Runs on the GPU box: how long the NVML queries of bench.py's clock sampler take while the GPU is busy, and what they cost a host-driven loop."""
import sys, time, threading
import numpy as np
sys.path.insert(0, ".")
import torch, pynvml
from bench import ITERS, make_problem
from momentum_b200 import solver as ms

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
ch, efs, theta0, _ = make_problem("cfg3-shard", 8192)
fn = ms.SkeletonSolverFunction(ch, 8192, efs, device=0); fn.upload_targets()
solver = ms.GaussNewtonSolver(ms.GaussNewtonSolverOptions(min_iterations=ITERS, max_iterations=ITERS, regularization=0.05), fn)
pins = [torch.from_numpy(theta0.astype(np.float32)).pin_memory() for _ in range(4)]
tps = [torch.from_numpy(np.ascontiguousarray(e.targets, np.float32)).pin_memory() for e in efs]
def step(i):
    for idx, tp in enumerate(tps): fn._check(fn._L.mb2_set_targets(fn._h, idx, ms.C.cast(tp.data_ptr(), ms._fp)))
    solver.solve_host_pointer(pins[i % 4].data_ptr()); solver.get_results()
for i in range(3): step(i)
def loop(n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): step(i)
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n
print("no sampler: %.3f ms per step" % loop())
for period in (0.005, 0.025, 0.05, 0.1):
    stop = threading.Event(); durs = {"clock": [], "reasons": []}
    def poll():
        while not stop.is_set():
            t0 = time.perf_counter(); pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); t1 = time.perf_counter()
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h); t2 = time.perf_counter()
            durs["clock"].append(t1 - t0); durs["reasons"].append(t2 - t1); time.sleep(period)
    th = threading.Thread(target=poll, daemon=True); th.start()
    ms_ = loop(); stop.set(); th.join()
    print("poll every %3.0f ms: %.3f ms per step; clock query median %.3f ms max %.3f, reasons query median %.3f ms max %.3f (%d polls)" % (
        1e3 * period, ms_, 1e3 * np.median(durs["clock"]), 1e3 * max(durs["clock"]), 1e3 * np.median(durs["reasons"]), 1e3 * max(durs["reasons"]), len(durs["clock"])))

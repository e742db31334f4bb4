#!/usr/bin/env python
"""Runs on the GPU box: the 64-instance mixed batch of tests/test_mixed_batch.py solved on every device path, each instance compared with
the float AND the double oracle. Prints per path how close the CUDA result is to the exact (double) answer next to how close the
reference's own float build is - the figure that says whether a path is noisier than the reference, independent of any threshold."""
import sys

import numpy as np

sys.path.insert(0, ".")
from momentum_b200 import character as mc, solver as ms  # noqa: E402
from momentum_b200.problems import mixed_problem  # noqa: E402
from oracle.binding import OracleFunction  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rigs, inst = mixed_problem(N, seed=11)
kw = dict(min_iterations=8, max_iterations=8, threshold=1.0, regularization=0.05)
ref32, ref64 = [], []
for x in inst:
    ch = rigs[x["rig"]][0]
    ef = mc.PositionErrorFunction(x["parents"], x["offsets"], x["weights"], x["targets"][None], weight=1.0)
    e32, p32, _, _ = OracleFunction(ch, [ef], "float32").solve(x["theta0"].astype(np.float64), **kw)
    e64, p64, _, _ = OracleFunction(ch, [ef], "float64").solve(x["theta0"].astype(np.float64), **kw)
    ref32.append((e32, p32)); ref64.append((e64, p64))


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))


g32 = np.array([rel(ref32[i][1], ref64[i][1]) for i in range(N)])
print(f"reference float vs double: median {np.median(g32):.2e}  p90 {np.quantile(g32, 0.9):.2e}  max {g32.max():.2e}  (> 1e-4: {(g32 > 1e-4).sum()}, > 2e-4: {(g32 > 2e-4).sum()})")
paths = [("auto", dict()), ("three kernels", dict(fused_mode=ms.FUSED_OFF)), ("gram+cholesky", dict(fused_mode=ms.FUSED_GRAM_CHOLESKY)), ("persistent", dict(fused_mode=ms.FUSED_PERSISTENT)),
         ("fp32 simt + eigen llt", dict(fused_mode=ms.FUSED_OFF, jtj_mode=ms.JTJ_FP32_SIMT, cholesky_mode=ms.CHOLESKY_DENSE_EIGEN))]
for name, extra in paths:
    mb = ms.MixedBatch()
    rid = {n: mb.add_rig(ch) for n, (ch, _) in rigs.items()}
    for x in inst:
        mb.add_instance(rid[x["rig"]], x["parents"], x["offsets"], x["weights"], x["targets"], x["theta0"])
    try:
        out = mb.solve(ms.GaussNewtonSolverOptions(**kw, **extra))
    except Exception as e:  # noqa: BLE001
        print(f"{name}: {e}")
        continue
    d32 = np.array([rel(out["params"][i], ref32[i][1]) for i in range(N)])
    d64 = np.array([rel(out["params"][i], ref64[i][1]) for i in range(N)])
    eabs = np.array([abs(out["errors"][i] - ref32[i][0]) / (abs(ref32[i][0]) + 1e-4) for i in range(N)])
    worse = int((d64 > np.maximum(1e-4, 2 * g32)).sum())
    print(f"{name:24s} d(cuda,f64): median {np.median(d64):.2e} p90 {np.quantile(d64, 0.9):.2e} max {d64.max():.2e} | d(cuda,f32): median {np.median(d32):.2e} "
          f"p90 {np.quantile(d32, 0.9):.2e} max {d32.max():.2e} (> 1e-4: {(d32 > 1e-4).sum()}, > 2e-4: {(d32 > 2e-4).sum()}) | farther from exact than max(1e-4, 2 x reference): {worse} | "
          f"objective rel diff max {eabs.max():.2e} status!=0: {(out['status'] != 0).sum()}")
    if name == "auto":
        for i in np.argsort(-d32)[:10]:
            print(f"     i={i} {inst[i]['rig']:12s} c={len(inst[i]['parents']):3d} d(cuda,f32)={d32[i]:.2e} d(cuda,f64)={d64[i]:.2e} d(f32,f64)={g32[i]:.2e} err {out['errors'][i]:.6g} / {ref32[i][0]:.6g} / {ref64[i][0]:.6g}")

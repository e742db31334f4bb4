#!/usr/bin/env python
"""Runs on the GPU box: how many instances of cfg2 / cfg3 / cfg4 miss the stated 1e-4 against the float oracle, and how those
instances compare with the oracle's own float-vs-double gap (the calibration tests/parity.py applies)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from momentum_b200 import solver as ms
from momentum_b200.problems import bodyhands_problem, humanoid_problem
from oracle.binding import OracleFunction
from tests import parity

def survey(name, ch, efs, theta0, opts):
    B = theta0.shape[0]
    fn = parity.build_function(ch, efs, B)
    out = ms.GaussNewtonSolver(opts, fn).solve(theta0)
    kw = dict(min_iterations=opts.min_iterations, max_iterations=opts.max_iterations, threshold=opts.threshold, regularization=opts.regularization)
    miss = []
    worst = 0.0
    for b in range(B):
        err, p, it, _ = OracleFunction(ch, efs, "float32", instance=b).solve(parity.f32(theta0[b]), **kw)
        d = np.max(np.abs(out["params"][b] - p)) / max(1.0, np.max(np.abs(p)))
        if d > 1e-4 or abs(out["errors"][b] - err) > 1e-3 * abs(err) + 1e-7:
            e64, p64, it64, _ = OracleFunction(ch, efs, "float64", instance=b).solve(parity.f32(theta0[b]), **kw)
            gap = np.max(np.abs(p - p64)) / max(1.0, np.max(np.abs(p)))
            dg = np.max(np.abs(out["params"][b] - p64)) / max(1.0, np.max(np.abs(p64)))
            miss.append((b, float(d), float(gap), float(dg), int(out["iterations"][b]), it, it64, float(out["errors"][b]), err, e64))
        else:
            worst = max(worst, d)
    print(f"{name}: {B} instances, {len(miss)} miss 1e-4 (worst of the others {worst:.2e})")
    for m in miss:
        print("   b=%d d(gpu,f32)=%.2e d(f32,f64)=%.2e d(gpu,f64)=%.2e its gpu/f32/f64=%d/%d/%d err gpu/f32/f64=%.6g/%.6g/%.6g" % m)

opts = ms.GaussNewtonSolverOptions(min_iterations=1, max_iterations=50, threshold=1.0, regularization=0.05)
for ori in (False, True):
    ch, efs, theta0, _ = humanoid_problem(96, orientation=ori)
    survey("cfg3" if ori else "cfg2", ch, efs, theta0, opts)
ch, efs, theta0, _ = bodyhands_problem(32)
survey("cfg4", ch, efs, theta0, opts)

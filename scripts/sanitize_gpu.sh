#!/bin/bash
# Runs on the GPU box: compute-sanitizer memcheck / racecheck / synccheck over one small solve of every kernel family
# (three kernels, Gram + Cholesky, persistent, line search, QR step, dense tcgen05 JtJ). Summaries -> gpurun_out/<tag>_san_<tool>.log
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/san_driver.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, ".")
from momentum_b200 import solver as ms
from momentum_b200.problems import humanoid_problem, chain_problem
from tests import parity
ch, efs, theta0, _ = humanoid_problem(6, orientation=True)
fn = parity.build_function(ch, efs, 6)
for kw in (dict(fused_mode=ms.FUSED_OFF), dict(fused_mode=ms.FUSED_GRAM_CHOLESKY), dict(fused_mode=ms.FUSED_PERSISTENT), dict(do_line_search=True, fused_mode=ms.FUSED_OFF),
           dict(linear_solver=ms.LINEAR_SOLVER_QR), dict(jtj_mode=ms.JTJ_TF32X3, cholesky_mode=ms.CHOLESKY_TILES_SPARSE, fused_mode=ms.FUSED_OFF),
           dict(cholesky_mode=ms.CHOLESKY_DENSE_EIGEN, fused_mode=ms.FUSED_OFF)):
    out = ms.GaussNewtonSolver(ms.GaussNewtonSolverOptions(min_iterations=2, max_iterations=2, regularization=0.05, **kw), fn).solve(theta0)
    print(kw, "status", out["status"].tolist(), "err0", float(out["errors"][0]))
ch, efs, theta0, _ = chain_problem(J=6, B=3, seed=21)
fn = parity.build_function(ch, efs, 3)
print(fn.get_error(theta0)[:2], ms.GaussNewtonSolver(ms.GaussNewtonSolverOptions(max_iterations=2), fn).solve(theta0)["errors"][:2])
# round-2 additions: the trust-region QR iteration, the wide (n = 424) tcgen05 JtJ with far items + the 16-warp Gram / Cholesky kernels of cfg4
print("trust region", ms.GaussNewtonSolver(ms.GaussNewtonSolverOptions(min_iterations=2, max_iterations=2, linear_solver=ms.LINEAR_SOLVER_TRUST_REGION_QR), fn).solve(theta0)["errors"][:2])
from momentum_b200.problems import bodyhands_problem
ch, efs, theta0, _ = bodyhands_problem(2)
fn = parity.build_function(ch, efs, 2)
for kw in (dict(), dict(jtj_mode=ms.JTJ_TF32X3, cholesky_mode=ms.CHOLESKY_TILES_SPARSE, fused_mode=ms.FUSED_OFF)):
    out = ms.GaussNewtonSolver(ms.GaussNewtonSolverOptions(min_iterations=1, max_iterations=1, regularization=0.05, **kw), fn).solve(theta0)
    print("bodyhands", kw, "status", out["status"].tolist(), "err0", float(out["errors"][0]))
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_driver.py > $OUT/${TAG}_san_${tool}.log 2>&1
  echo "== $tool: $(grep -c 'ERROR SUMMARY\|RACECHECK SUMMARY' $OUT/${TAG}_san_${tool}.log) summary lines"; grep "SUMMARY" $OUT/${TAG}_san_${tool}.log | tail -3
done

import sys
import numpy as np
sys.path.insert(0, ".")
from momentum_b200 import solver as ms
from momentum_b200.problems import humanoid_problem
from tests import parity
ch, efs, theta0, _ = humanoid_problem(16, orientation=True)
fn = parity.build_function(ch, efs, 16)
for nit in (1, 2, 3, 10):
    outs = []
    for fm in (ms.FUSED_OFF, ms.FUSED_ON):
        s = ms.GaussNewtonSolver(ms.GaussNewtonSolverOptions(regularization=0.05, fused_mode=fm, store_error_history=True, min_iterations=nit, max_iterations=nit), fn)
        o = s.solve(theta0); o["hist"] = s.get_error_history(); outs.append(o)
    a, b = outs
    dp = np.abs(a["params"] - b["params"])
    print("iters", nit, "max param diff", dp.max(), "argmax", np.unravel_index(dp.argmax(), dp.shape), "hist rel diff per it", np.max(np.abs(a["hist"] - b["hist"]) / np.abs(a["hist"]), axis=0))
    if nit == 1:
        print(" step (unfused) max", np.abs(a["params"] - theta0).max(), "per-instance max diff", dp.max(axis=1))
        bad = np.argwhere(dp > 1e-6)
        print(" entries > 1e-6:", bad[:40].tolist())

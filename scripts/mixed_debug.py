import sys
import numpy as np
sys.path.insert(0, ".")
from momentum_b200 import character as mc, solver as ms
from momentum_b200.problems import mixed_problem
from oracle.binding import OracleFunction
rigs, inst = mixed_problem(64, seed=11)
mb = ms.MixedBatch()
rid = {name: mb.add_rig(ch) for name, (ch, _) in rigs.items()}
for x in inst:
    mb.add_instance(rid[x["rig"]], x["parents"], x["offsets"], x["weights"], x["targets"], x["theta0"])
out = mb.solve(ms.GaussNewtonSolverOptions(min_iterations=8, max_iterations=8, threshold=1.0, regularization=0.05))
for i, x in enumerate(inst):
    if out["status"][i] != 0 or i < 4:
        ch = rigs[x["rig"]][0]
        ef = mc.PositionErrorFunction(x["parents"], x["offsets"], x["weights"], x["targets"][None], weight=1.0)
        err, p, it, _ = OracleFunction(ch, [ef], "float32").solve(x["theta0"].astype(np.float64), min_iterations=8, max_iterations=8, threshold=1.0, regularization=0.05)
        d = np.max(np.abs(out["params"][i] - p)) / max(1.0, np.max(np.abs(p)))
        print(i, x["rig"], "c", len(x["parents"]), "status", out["status"][i], "err gpu/orc", out["errors"][i], err, "d", d, "finite orc", np.isfinite(p).all(), "max|p|", np.abs(p).max())

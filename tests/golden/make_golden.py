#!/usr/bin/env python
"""Generates tests/golden/*.npz: outputs of the double-precision oracle (oracle/ik_oracle.hpp, itself pinned to the
reference's known answers by tests/test_oracle_known_answers.py) on small seeded problems. The reference is C++ on Eigen 5 and
cannot be built or imported in this image, so the fixtures come from its restatement; they freeze the oracle (a CPU test
re-derives them) and give the GPU parity tests a comparison that does not depend on the oracle build on the GPU box.

    python tests/golden/make_golden.py          # rewrites the fixtures
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from momentum_b200.problems import chain_problem, humanoid_problem  # noqa: E402
from oracle.binding import OracleFunction  # noqa: E402

CASES = {
    # name: (problem factory, instances, solver options)
    "chain6_all_families": (lambda: chain_problem(J=6, B=3, seed=25), dict(min_iterations=1, max_iterations=12, threshold=10.0, regularization=0.05)),
    "humanoid72_cfg3": (lambda: humanoid_problem(3, orientation=True), dict(min_iterations=1, max_iterations=8, threshold=1.0, regularization=0.05)),
}


def generate(name):
    make, opts = CASES[name]
    ch, efs, theta0, _ = make()
    B = theta0.shape[0]
    out = {"theta0": theta0.astype(np.float32)}
    errs, jtr, jtj_diag, sol, sol_err, its, fk = [], [], [], [], [], [], []
    for b in range(B):
        orc = OracleFunction(ch, efs, "float64", instance=b)
        th = np.asarray(theta0[b], np.float32).astype(np.float64)
        errs.append(orc.get_error(th))
        _, H, g = orc.get_jtjr(th)
        jtr.append(g)
        jtj_diag.append(np.diag(H).copy())
        xf, _, _ = orc.fk(th)
        fk.append(xf)
        e, p, it, _ = orc.solve(th, **opts)
        sol.append(p); sol_err.append(e); its.append(it)
    out.update(error=np.array(errs), jtr=np.array(jtr), jtj_diag=np.array(jtj_diag), fk=np.array(fk), solution=np.array(sol),
               solution_error=np.array(sol_err), iterations=np.array(its), options=np.array([opts[k] for k in ("min_iterations", "max_iterations", "threshold", "regularization")], np.float64))
    return out


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"), **generate(name))
        print("wrote", name)

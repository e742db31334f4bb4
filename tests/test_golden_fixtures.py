"""tests/golden/*.npz (made by tests/golden/make_golden.py from the double-precision oracle): the oracle must keep reproducing
them (CPU), and the CUDA path must match them through the C-ABI (GPU) -- a parity check that needs no oracle build on the GPU box."""
import importlib.util
import os

import numpy as np
import pytest

from momentum_b200 import solver as ms
from tests import parity

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden_fixtures(name):
    gold, now = load(name), make_golden.generate(name)
    for key in gold.files:
        np.testing.assert_allclose(now[key], gold[key], rtol=1e-9, atol=1e-12, err_msg=f"{name}:{key}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_cuda_path_matches_golden_fixtures(name):
    gold = load(name)
    make, opts = make_golden.CASES[name]
    ch, efs, theta0, _ = make()
    B = theta0.shape[0]
    assert np.array_equal(theta0.astype(np.float32), gold["theta0"])
    fn = parity.build_function(ch, efs, B)
    err = fn.get_error(theta0)
    np.testing.assert_allclose(err, gold["error"], rtol=2e-5, atol=1e-7)
    _, H, g = fn.get_jtjr(theta0, ms.JTJ_FP32_SIMT)
    scale = np.maximum(1.0, np.abs(gold["jtr"]).max(axis=1, keepdims=True))
    assert np.max(np.abs(g - gold["jtr"]) / scale) <= 2e-5
    d = np.stack([np.diag(H[b]) for b in range(B)])
    assert np.max(np.abs(d - gold["jtj_diag"]) / np.maximum(1.0, np.abs(gold["jtj_diag"]).max(axis=1, keepdims=True))) <= 2e-5
    st = fn.get_skeleton_state(theta0)
    assert np.max(np.abs(st[:, :, :3] - gold["fk"][:, :, :3])) <= 2e-5 * max(1.0, np.abs(gold["fk"][:, :, :3]).max())
    solver = ms.GaussNewtonSolver(ms.GaussNewtonSolverOptions(**opts), fn)
    out = solver.solve(theta0)
    # float path against the double fixture: the reference's own float-vs-double bound (error_function_helpers.h:38-52) is 5e-3 / 1e-3 ...
    for b in range(B):
        p = gold["solution"][b]
        rel = np.max(np.abs(out["params"][b] - p)) / max(1.0, np.max(np.abs(p)))
        assert rel <= 1e-3, (name, b, rel)
        assert abs(out["errors"][b] - gold["solution_error"][b]) <= 2e-3 * abs(gold["solution_error"][b]) + 1e-6

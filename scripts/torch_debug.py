import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch
from momentum_b200 import torch_ik as ti, character as mc
from oracle.binding import OracleFunction
from test_torch_ik import _problem
ch, parents, offsets, targets, active, _ = _problem(B=2, seed=9)
B, n = targets.shape[0], ch.num_params
dev = torch.device("cuda", 0)
rng = np.random.default_rng(1)
gout = torch.from_numpy(rng.normal(size=(B, n))).to(dev).float()
pw0 = (1.0 + 0.3 * torch.rand(B, len(parents), device=dev, dtype=torch.float64))
for (lam, its, ls) in ((1e-4, 40, False), (0.01, 200, False), (0.01, 60, True)):
    opts = ti.SolverOptions(levmar_lambda=lam, min_iter=its, max_iter=its, threshold=1.0, line_search=ls)
    def run(tg, efw, pw):
        return ti.solve_ik(ch, active, torch.zeros(B, n, device=dev), [ti.ErrorFunctionType.Position], efw, opts, position_cons_parents=parents, position_cons_offsets=offsets,
                           position_cons_weights=pw, position_cons_targets=tg)
    tg = torch.from_numpy(targets).to(dev).double().requires_grad_(True)
    efw = torch.ones(B, 1, device=dev, dtype=torch.float64, requires_grad=True)
    pw = pw0.clone().requires_grad_(True)
    theta = run(tg, efw, pw)
    (theta.float() * gout).sum().backward()
    th = theta.detach().cpu().numpy()
    print("lambda", lam, "its", its, "ls", ls, "errors", ti.solve_ik.last_results["errors"])
    # oracle IFT in double at theta_final of instance 0
    b = 0
    ef = mc.PositionErrorFunction(parents, offsets, pw0[b].cpu().numpy(), targets, weight=1.0)
    orc = OracleFunction(ch, [ef], "float64", instance=b)
    e, J, r, rows = orc.get_jacobian(th[b].astype(np.float64))
    J = J[:24]; r = r[:24]
    act = np.nonzero(active)[0]
    Ja = J[:, act]
    grad = 2 * Ja.T @ r
    print("  grad rms", np.sqrt((grad ** 2).mean()), "objective", e)
    U, S, Vt = np.linalg.svd(Ja, full_matrices=False)
    g = gout[b].cpu().numpy().astype(np.float64)[act]
    tmp = Vt @ g; tmp = np.where(S * S < 1e-5, 0, tmp / (S * S)); va = 0.5 * Vt.T @ tmp
    v = np.zeros(n); v[act] = va
    w = pw0[b].cpu().numpy()
    gt = 2 * np.sqrt(w)[:, None] * (J @ v).reshape(-1, 3)
    print("  oracle IFT dL/dt[0,0,:]", gt[0], " torch", tg.grad[b, 0].cpu().numpy())
    with torch.no_grad():
        for eps in (1e-2, 2e-3):
            d = torch.zeros_like(tg); d[0, 0, 0] = eps
            fd = ((run(tg + d, efw, pw).float() * gout).sum() - (run(tg - d, efw, pw).float() * gout).sum()) / (2 * eps)
            print("  FD eps", eps, fd.item())

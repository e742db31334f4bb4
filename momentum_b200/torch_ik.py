"""Torch front end of the batched device solver: ``solve_ik`` on CUDA tensors, with the implicit-function backward.

Mirror of ``pymomentum.solver.solve_ik`` (pymomentum/tensor_ik/tensor_ik.cpp:95-188: per batch element build the error functions, a
solver function, a solver, solve, NaN guard) for the error functions of the device path — Limit, Position, Orientation, Motion
(= ModelParametersErrorFunction) — with the same keyword names. The forward pass never leaves the GPU: targets and weights are read
from the caller's tensors (``mb2_set_targets_device`` / ``mb2_set_constraint_weights_device``), the solve runs on torch's current
stream (``mb2_solver_solve_device``) and the result is a CUDA tensor.

The backward pass is ``d_solveTensorIKProblem`` (tensor_ik.cpp:191-340) with ``d_modelParams_d_inputs``
(momentum/diff_ik/fully_differentiable_body_ik.cpp:112-238): at the solution, v = (2 J^T J)^+ dLoss/dtheta through the SVD of the
Jacobian restricted to the active parameters (singular values with s^2 < 1e-5 dropped), then
    dLoss/d weight_k       = -(grad_theta E_k / weight_k) . v
    dLoss/d input          = d/d input [ grad_theta E_k . (-v) ]
and no gradient for an element whose gradient RMS exceeds 0.01 (the solve did not converge, tensor_ik.cpp:254). The Jacobian comes
from the device (``mb2_solver_function_get_jacobian_device``); the small dense algebra is torch on the same device.
torch is plumbing here (tensors, streams, autograd bookkeeping): every kernel on the forward path is this repo's.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import IntEnum
from typing import Optional, Sequence

import numpy as np
import torch

from . import character as mc
from . import solver as ms

GRADIENT_RMSE_THRESHOLD = 0.01  # tensor_ik.cpp:46


class ErrorFunctionType(IntEnum):
    """The subset of pymomentum's ErrorFunctionType served by the device path."""

    Position = 0
    Orientation = 1
    Limit = 2
    Motion = 3


class LinearSolverType(IntEnum):
    """pymomentum/tensor_ik/solver_options.h:12-26."""

    Cholesky = 0       # SubsetGaussNewtonSolverT (tensor_ik.cpp:143-148): the tile-scheduled device path (the fast one)
    QR = 1             # GaussNewtonSolverQRT (tensor_ik.cpp:153-158): Householder QR of [sqrt(lambda) I; J], the reference's default
    TrustRegionQR = 2  # TrustRegionQRT (tensor_ik.cpp:149-152): the device trust-region iteration (ik_tr_qr.cuh); levmar_lambda / line_search do not apply


@dataclass
class SolverOptions:
    """pymomentum/tensor_ik/solver_options.h:27-47. The reference defaults to QR; here the default is Cholesky, the path this
    repository accelerates (same normal equations, same minimiser); QR and TrustRegionQR run the device Householder kernels."""

    linear_solver_type: LinearSolverType = LinearSolverType.Cholesky
    levmar_lambda: float = 0.01
    min_iter: int = 4
    max_iter: int = 50
    threshold: float = 10.0
    line_search: bool = True
    verbose: bool = False


_handles = {}


def _f32c(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _build(character: mc.Character, B, device_index, pos_parents, pos_offsets, ori_parents, ori_offsets, motion_weights, use_limit, active):
    """One cached solver function per (character, batch, constraint topology): the plan is built once."""
    key = (id(character), B, device_index, None if pos_parents is None else pos_parents.tobytes(), None if pos_offsets is None else pos_offsets.tobytes(),
           None if ori_parents is None else ori_parents.tobytes(), None if ori_offsets is None else ori_offsets.tobytes(),
           None if motion_weights is None else motion_weights.tobytes(), use_limit, active.tobytes())
    if key in _handles:
        return _handles[key]
    fn = ms.SkeletonSolverFunction(character, B, device=device_index)
    blocks = {}
    n = character.num_params
    if pos_parents is not None:
        nc = len(pos_parents)
        blocks["position"] = fn.add_error_function(mc.PositionErrorFunction(pos_parents, pos_offsets, np.ones(nc, np.float32), np.zeros((B, nc, 3), np.float32), weight=1.0))
    if ori_parents is not None:
        nc = len(ori_parents)
        blocks["orientation"] = fn.add_error_function(mc.OrientationErrorFunction(ori_parents, ori_offsets, np.ones(nc, np.float32), np.zeros((B, nc, 4), np.float32), weight=1.0))
    if use_limit:
        blocks["limit"] = fn.add_error_function(mc.LimitErrorFunction(weight=1.0))
    if motion_weights is not None:
        blocks["motion"] = fn.add_error_function(mc.ModelParametersErrorFunction(motion_weights, np.zeros((B, n), np.float32), weight=1.0))
    fn.set_enabled_parameters(active)
    _handles[key] = (fn, blocks)
    return _handles[key]


class _SolveIK(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, theta0, efw, pos_targets, pos_weights, ori_targets, ori_weights, motion_targets):
        fn, blocks, opts, kinds = cfg["fn"], cfg["blocks"], cfg["options"], cfg["kinds"]
        dev = theta0.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        B = theta0.shape[0]
        efw32 = _f32c(efw, dev)
        keep = []  # device buffers the asynchronous calls read
        # per-element error-function weights (buildMomentumErrorFunctions, tensor_ik_utility.cpp:149-181): folded into the per-instance
        # constraint weights for Position / Orientation; Limit and Motion carry one weight for the whole batch on the device
        for name, tgt, w in (("position", pos_targets, pos_weights), ("orientation", ori_targets, ori_weights)):
            if name not in blocks:
                continue
            k = kinds.index(ErrorFunctionType.Position if name == "position" else ErrorFunctionType.Orientation)
            t32 = _f32c(tgt, dev)
            w32 = (_f32c(w, dev) * efw32[:, k:k + 1]).contiguous()
            fn.set_targets_device(blocks[name], t32.data_ptr(), stream)
            fn.set_constraint_weights_device(blocks[name], w32.data_ptr(), stream)
            keep += [t32, w32]
        for name, kind in (("limit", ErrorFunctionType.Limit), ("motion", ErrorFunctionType.Motion)):
            if name not in blocks:
                continue
            col = efw32[:, kinds.index(kind)]
            w0 = float(col[0])
            if not bool(torch.all(col == col[0])):
                raise ValueError(f"{name} error-function weights must be the same for every batch element on the device path")
            fn.set_error_function_weight(blocks[name], w0)
        if "motion" in blocks:
            m32 = _f32c(motion_targets, dev)
            fn.set_targets_device(blocks["motion"], m32.data_ptr(), stream)
            keep.append(m32)
        solver = ms.GaussNewtonSolver(opts, fn)
        theta = _f32c(theta0, dev).clone()
        solver.solve_device(theta.data_ptr(), stream)
        res = solver.get_results()  # synchronises; the NaN / Inf guard (tensor_ik.cpp:168-173) already ran on the device
        ctx.cfg = cfg
        ctx.status = res["status"]
        ctx.save_for_backward(theta, efw32, *(keep))
        ctx.keep_names = [n for n in ("position", "orientation") if n in blocks]
        ctx.in_dtypes = (efw.dtype, None if pos_targets is None else pos_targets.dtype, None if pos_weights is None else pos_weights.dtype)
        cfg["last_results"] = res
        return theta.to(theta0.dtype)

    @staticmethod
    def backward(ctx, grad_theta):
        cfg = ctx.cfg
        fn, blocks, kinds, active = cfg["fn"], cfg["blocks"], cfg["kinds"], cfg["active"]
        saved = ctx.saved_tensors
        theta, efw32 = saved[0], saved[1]
        dev = theta.device
        B, n = theta.shape
        stream = torch.cuda.current_stream(dev).cuda_stream
        ptr, ld = fn.get_jacobian_device(theta.data_ptr(), stream)
        rows = sum(cfg["block_rows"].values())
        torch.cuda.current_stream(dev).synchronize()
        J_all = _device_view(ptr, (B, n + 1, ld), dev).clone()  # the handle's own buffer [B][n + 1][ld], copied before the next call reuses it
        J = J_all[:, :n, :rows].transpose(1, 2).double()   # [B, rows, n]
        r = J_all[:, n, :rows].double()                    # [B, rows]
        act = torch.as_tensor(np.nonzero(active)[0], device=dev)
        Ja = J[:, :, act]
        g = grad_theta.to(dev).double()[:, act]
        grad_rms = torch.sqrt(((2.0 * torch.einsum("brk,br->bk", Ja, r)) ** 2).mean(dim=1))
        ok = (grad_rms <= GRADIENT_RMSE_THRESHOLD).double()[:, None]
        # v = (2 J^T J)^+ g through the SVD of J (fully_differentiable_body_ik.cpp:78-109)
        U, S, Vh = torch.linalg.svd(Ja, full_matrices=False)
        s2 = S * S
        tmp = torch.einsum("bkn,bn->bk", Vh, g)
        tmp = torch.where(s2 < 1e-5, torch.zeros_like(tmp), tmp / s2.clamp_min(1e-30))
        va = 0.5 * torch.einsum("bkn,bk->bn", Vh, tmp)
        v = torch.zeros(B, n, dtype=torch.float64, device=dev)
        v[:, act] = va
        Jv = torch.einsum("brn,bn->br", J, v) * ok          # [B, rows], zero for unconverged elements
        # rows of each block in the API-parity layout: blocks in the order they were added
        grad_efw = torch.zeros(B, len(kinds), dtype=torch.float64, device=dev)
        grads = {"pos_targets": None, "pos_weights": None}
        row = 0
        sizes = cfg["block_rows"]
        idx = 2
        for name in cfg["block_order"]:
            nr = sizes[name]
            rk, Jvk = r[:, row:row + nr], Jv[:, row:row + nr]
            kind = {"position": ErrorFunctionType.Position, "orientation": ErrorFunctionType.Orientation, "limit": ErrorFunctionType.Limit, "motion": ErrorFunctionType.Motion}[name]
            k = kinds.index(kind)
            wk = efw32[:, k].double()
            # E_k = sum r^2, grad E_k = 2 J_k^T r_k (rows already carry sqrt(weight)): dLoss/dw_k = -(grad E_k / w_k) . v
            grad_efw[:, k] = torch.where(wk != 0, -2.0 * (rk * Jvk).sum(dim=1) / wk.clamp_min(1e-30) * (wk != 0), torch.zeros_like(wk))
            if name == "position":
                nc = nr // 3
                w_eff = saved[idx + 1].double()             # [B, nc] = constraint weight x error-function weight
                sq = torch.sqrt(w_eff)[:, :, None]
                Jv3, r3 = Jvk.reshape(B, nc, 3), rk.reshape(B, nc, 3)
                # rows = sqrt(w) (p - t): d/dt [grad E . (-v)] = 2 sqrt(w) (J_dev v)
                grads["pos_targets"] = 2.0 * sq * Jv3
                # d/d(constraint weight): grad E_c = 2 w J_c^T f_c -> -(2 J_c^T f_c) . v * efw = -(2 r_c . Jv_c) / w_eff * efw
                cwg = -2.0 * (r3 * Jv3).sum(dim=2) / w_eff.clamp_min(1e-30) * (w_eff != 0)
                grads["pos_weights"] = cwg * wk[:, None]
                idx += 2
            elif name == "orientation":
                idx += 2
            row += nr
        in_dt = ctx.in_dtypes
        g_efw = grad_efw.to(in_dt[0])
        g_pt = None if grads["pos_targets"] is None or in_dt[1] is None else grads["pos_targets"].to(in_dt[1])
        g_pw = None if grads["pos_weights"] is None or in_dt[2] is None else grads["pos_weights"].to(in_dt[2])
        return None, None, g_efw, g_pt, g_pw, None, None, None


def _device_view(ptr: int, shape, device):
    """float32 CUDA tensor over foreign device memory (the handle's Jacobian buffer) through __cuda_array_interface__."""

    class _Arr:
        pass

    a = _Arr()
    a.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(a, device=device)


def solve_ik(character: mc.Character, active_parameters, model_parameters_init: torch.Tensor, active_error_functions: Sequence[ErrorFunctionType],
             error_function_weights: torch.Tensor, options: Optional[SolverOptions] = None,
             position_cons_parents=None, position_cons_offsets=None, position_cons_weights=None, position_cons_targets=None,
             orientation_cons_parents=None, orientation_cons_offsets=None, orientation_cons_weights=None, orientation_cons_targets=None,
             motion_targets=None, motion_weights=None) -> torch.Tensor:
    """Batched IK on the GPU; arguments as pymomentum.solver.solve_ik. ``model_parameters_init`` [B, n] must live on a CUDA device;
    constraint parents / offsets are shared by the batch ([nc] / [nc, 3] / [nc, 4]), targets and weights are per element.
    Differentiable w.r.t. ``error_function_weights``, ``position_cons_targets`` and ``position_cons_weights``."""
    if not model_parameters_init.is_cuda:
        raise ValueError("momentum_b200.torch_ik.solve_ik runs on CUDA tensors (there is no CPU fallback)")
    options = options or SolverOptions()
    dev = model_parameters_init.device
    B, n = model_parameters_init.shape
    kinds = [ErrorFunctionType(k) for k in active_error_functions]
    efw = error_function_weights
    if efw.dim() == 1:
        efw = efw[None].expand(B, -1)
    active = np.asarray(active_parameters.cpu() if torch.is_tensor(active_parameters) else active_parameters, bool)

    def np_or_none(x, dt):
        return None if x is None else np.ascontiguousarray(x.detach().cpu().numpy() if torch.is_tensor(x) else x, dt)

    use_pos = ErrorFunctionType.Position in kinds and position_cons_parents is not None
    use_ori = ErrorFunctionType.Orientation in kinds and orientation_cons_parents is not None
    use_motion = ErrorFunctionType.Motion in kinds and motion_targets is not None
    pp = np_or_none(position_cons_parents, np.int32) if use_pos else None
    po = (np_or_none(position_cons_offsets, np.float32) if position_cons_offsets is not None else np.zeros((len(pp), 3), np.float32)) if use_pos else None
    op = np_or_none(orientation_cons_parents, np.int32) if use_ori else None
    oo = (np_or_none(orientation_cons_offsets, np.float32) if orientation_cons_offsets is not None else np.tile(np.array([0, 0, 0, 1], np.float32), (len(op), 1))) if use_ori else None
    mw = None
    if use_motion:
        mwt = motion_weights if motion_weights is not None else torch.ones(n)
        mw = np_or_none(mwt[0] if (torch.is_tensor(mwt) and mwt.dim() == 2) else mwt, np.float32)
    fn, blocks = _build(character, B, dev.index or 0, pp, po, op, oo, mw, ErrorFunctionType.Limit in kinds, active)
    order = [name for name in ("position", "orientation", "limit", "motion") if name in blocks]
    block_rows = {}
    for name in order:
        block_rows[name] = {"position": lambda: 3 * len(pp), "orientation": lambda: 9 * len(op),
                            "limit": lambda: mc.jacobian_size(character, mc.LimitErrorFunction()), "motion": lambda: int(((mw > 0) & active[: len(mw)]).sum())}[name]()
    lst = LinearSolverType(options.linear_solver_type)
    linear = {LinearSolverType.Cholesky: ms.LINEAR_SOLVER_CHOLESKY, LinearSolverType.QR: ms.LINEAR_SOLVER_QR, LinearSolverType.TrustRegionQR: ms.LINEAR_SOLVER_TRUST_REGION_QR}[lst]
    # Cholesky: SubsetGaussNewtonSolverT's line search (c1 = 1e-4 with the directional derivative); QR: GaussNewtonSolverQRT's (same rule,
    # gauss_newton_solver_qr.cpp:118-143): both are the subset variant on the device. TrustRegionQR takes the plain SolverOptions
    # (tensor_ik.cpp:149-152): iterations / threshold only, radius 1.
    opts = ms.GaussNewtonSolverOptions(min_iterations=options.min_iter, max_iterations=options.max_iter, threshold=options.threshold, regularization=options.levmar_lambda,
                                       do_line_search=options.line_search, subset_line_search=True, linear_solver=linear)
    cfg = {"fn": fn, "blocks": blocks, "options": opts, "kinds": kinds, "active": active, "block_order": order, "block_rows": block_rows}
    ones = lambda nc: torch.ones(B, nc, device=dev)
    pw = position_cons_weights if position_cons_weights is not None else (ones(len(pp)) if use_pos else None)
    ow = orientation_cons_weights if orientation_cons_weights is not None else (ones(len(op)) if use_ori else None)
    out = _SolveIK.apply(cfg, model_parameters_init, efw, position_cons_targets if use_pos else None, pw if use_pos else None,
                         orientation_cons_targets if use_ori else None, ow if use_ori else None, motion_targets if use_motion else None)
    solve_ik.last_results = cfg.get("last_results")
    return out

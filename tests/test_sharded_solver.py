"""Single-process multi-GPU front end (mb2_sharded_solver_*): a batch cut into contiguous blocks, one per device.
CPU: the host-side logic that needs no device (argument checks, no CPU fallback). GPU: shards against one unsharded solver."""
import ctypes as C
import dataclasses

import numpy as np
import pytest

from momentum_b200 import solver as ms
from momentum_b200.problems import humanoid_problem


def test_sharded_create_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = ms.load_library()
    h = C.c_void_p()
    dev = (C.c_int32 * 2)(0, 1)
    o = ms.GaussNewtonSolverOptions()._c()
    # null prototype: argument check comes first
    assert L.mb2_sharded_solver_create(None, 8, 2, dev, C.byref(o), C.byref(h)) != 0
    assert b"null" in L.mb2_sharded_last_error()
    assert L.mb2_sharded_solver_num_shards(None) == 0
    assert L.mb2_character_device(None) == -1 and L.mb2_solver_function_target_size(None, 0) == -1


def _solve_single(ch, efs, theta0, opts):
    fn = ms.SkeletonSolverFunction(ch, theta0.shape[0], efs)
    fn.upload_targets()
    return fn, ms.GaussNewtonSolver(opts, fn).solve(theta0)


@pytest.mark.gpu
@pytest.mark.parametrize("shards", [2, 3])
def test_shards_on_one_device_match_the_unsharded_solve_bit_for_bit(shards):
    """devices = [0, 0(, 0)]: the blocks are independent solves, so every instance's result is the one the whole batch gives."""
    B = 37  # not a multiple of the shard count: block sizes differ by one
    ch, efs, theta0, _ = humanoid_problem(B, orientation=True)
    opts = ms.GaussNewtonSolverOptions(min_iterations=1, max_iterations=12, threshold=1.0, regularization=0.05)
    fn, ref = _solve_single(ch, efs, theta0, opts)
    sh = ms.ShardedGaussNewtonSolver(opts, fn, B, [0] * shards, error_functions=efs)
    info = sh.shards()
    assert [s["count"] for s in info] == [B // shards + (1 if k < B % shards else 0) for k in range(shards)]
    assert info[0]["first"] == 0 and all(info[k]["first"] == info[k - 1]["first"] + info[k - 1]["count"] for k in range(1, shards))
    out = sh.solve(theta0)
    assert np.array_equal(out["params"], ref["params"]) and np.array_equal(out["iterations"], ref["iterations"]) and np.array_equal(out["status"], ref["status"])
    assert np.array_equal(out["errors"], ref["errors"])
    agg = out["aggregate"]
    assert agg["iterations"] == int(ref["iterations"].sum()) and agg["instances_ok"] == B
    assert agg["error_sum"] == pytest.approx(float(np.sum(ref["errors"])), rel=1e-12)
    # new targets for the whole batch, new options: the replicas follow
    efs2 = [dataclasses.replace(e, targets=np.asarray(e.targets)[::-1].copy()) for e in efs]
    for i, e in enumerate(efs2):
        sh.set_targets(i, e.targets)
    fn2, ref2 = _solve_single(ch, efs2, theta0, opts)
    out2 = sh.solve(theta0)
    assert np.array_equal(out2["params"], ref2["params"])


@pytest.mark.gpu
def test_every_visible_device_takes_a_shard():
    """One shard per visible GPU (1 on the single-GPU test box, N under gpurun --gpus N); results equal the unsharded solve."""
    L = ms.load_library()
    ndev = L.mb2_device_count()
    B = 16 * max(ndev, 1) + 3
    ch, efs, theta0, _ = humanoid_problem(B, orientation=False)
    opts = ms.GaussNewtonSolverOptions(min_iterations=6, max_iterations=6, regularization=0.05)
    fn, ref = _solve_single(ch, efs, theta0, opts)
    sh = ms.ShardedGaussNewtonSolver(opts, fn, B, list(range(ndev)), error_functions=efs)
    assert [s["device"] for s in sh.shards()] == list(range(ndev))
    out = sh.solve(theta0)
    assert np.array_equal(out["params"], ref["params"]) and out["aggregate"]["iterations"] == 6 * B
    with pytest.raises(ms.MomentumB200Error, match="not a usable"):
        ms.ShardedGaussNewtonSolver(opts, fn, B, [0, ndev + 5])

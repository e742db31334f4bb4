"""Seeded synthetic problems shared by the oracle tests, the GPU parity tests and bench.py."""
import numpy as np

from momentum_b200 import character as mc


def random_unit_quats(rng, shape):
    q = rng.normal(size=tuple(shape) + (4,))
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def small_rotation_quats(rng, shape, max_angle):
    ax = rng.normal(size=tuple(shape) + (3,))
    ax /= np.linalg.norm(ax, axis=-1, keepdims=True)
    ang = rng.uniform(0, max_angle, size=tuple(shape) + (1,))
    return np.concatenate([ax * np.sin(ang / 2), np.cos(ang / 2)], -1)


def add_test_limits(ch, rng, ellipsoid=True):
    """A limit of every live type (character/parameter_limits.h:20-33) on a chain test character."""
    n, J = ch.num_params, ch.num_joints
    L = [mc.ParameterLimit(mc.LIMIT_MINMAX, 1.0, (0,), (-0.1, 0.1)),
         mc.ParameterLimit(mc.LIMIT_MINMAX, 0.7, (3,), (-0.2, 0.15)),
         mc.ParameterLimit(mc.LIMIT_MINMAX_JOINT, 1.3, (1, 3), (-0.1, 0.1)),
         mc.ParameterLimit(mc.LIMIT_MINMAX_JOINT, 0.9, (2, 5), (-0.05, 0.05)),
         mc.ParameterLimit(mc.LIMIT_MINMAX_JOINT_PASSIVE, 1.0, (1, 3), (-0.1, 0.1)),
         mc.ParameterLimit(mc.LIMIT_LINEAR, 0.8, (7, 8), (0.5, 0.1, 0.0, 0.0)),
         mc.ParameterLimit(mc.LIMIT_LINEAR, 0.6, (3, 4), (1.5, -0.1, -10.0, 10.0)),
         mc.ParameterLimit(mc.LIMIT_LINEAR_JOINT, 1.1, (1, 3, 2, 3), (0.7, 0.05, 0.0, 0.0)),
         mc.ParameterLimit(mc.LIMIT_HALFPLANE, 1.2, (3, 5), (0.6, 0.8, 0.3)),
         mc.ParameterLimit(mc.LIMIT_HALFPLANE, 1.0, (7, 9), (-0.8, 0.6, 0.2))]
    A = np.diag(rng.uniform(0.5, 1.5, 3)) @ np.linalg.qr(rng.normal(size=(3, 3)))[0]
    t = rng.uniform(-0.2, 0.2, 3)
    M = np.concatenate([A, t[:, None]], 1)
    Ai = np.linalg.inv(A)
    Mi = np.concatenate([Ai, (-Ai @ t)[:, None]], 1)
    off = rng.uniform(-0.5, 0.5, 3)
    if ellipsoid:
        L.append(mc.ParameterLimit(mc.LIMIT_ELLIPSOID, 0.9, (1, min(4, J - 1)), tuple(M.reshape(-1)) + tuple(Mi.reshape(-1)) + tuple(off)))
    ch.limits = L
    return ch


def chain_problem(J=6, B=3, seed=0, families=("position", "orientation", "state", "limit"), loss=(mc.LOSS_L2, 1.0), logmap=False,
                  rot_diff=False, ellipsoid=True):
    """Small all-families problem on the reference test fixture; targets = perturbed reachable poses."""
    rng = np.random.default_rng(seed)
    ch = mc.create_test_character(J)
    add_test_limits(ch, rng, ellipsoid)
    n = ch.num_params
    theta_star = rng.uniform(-0.5, 0.5, (B, n))
    efs = []
    if "position" in families:
        par = rng.integers(0, J, 5).astype(np.int32)
        off = rng.uniform(-1, 1, (5, 3))
        w = rng.uniform(0.5, 1.5, 5); w[3] = 0.0  # zero-weight constraint is skipped (joint_error_function-inl.h:197)
        tg = mc.world_points(ch, theta_star, par, off) + 0.05 * rng.normal(size=(B, 5, 3))
        efs.append(mc.PositionErrorFunction(par, off, w, tg, weight=0.9, loss_alpha=loss[0], loss_c=loss[1]))
    if "orientation" in families:
        par = rng.integers(0, J, 3).astype(np.int32)
        off = random_unit_quats(rng, (3,))
        tg = mc._qmul(mc.world_rotations(ch, theta_star, par, off), small_rotation_quats(rng, (B, 3), 0.3))
        efs.append(mc.OrientationErrorFunction(par, off, rng.uniform(0.5, 1.5, 3), tg, weight=0.4, loss_alpha=loss[0], loss_c=loss[1], rot_diff=rot_diff))
    if "state" in families:
        t, q, s = mc.forward_kinematics(ch, theta_star)
        tg = np.concatenate([t + 0.05 * rng.normal(size=t.shape), mc._qmul(q, small_rotation_quats(rng, (B, J), 0.3)), s[..., None]], -1)
        pw = rng.uniform(0.5, 1.5, J); rw = rng.uniform(0.5, 1.5, J)
        pw[1] = 0; rw[1] = 0  # inactive joint contributes no rows (state_error_function.cpp:424-426)
        if J > 3:
            pw[3] = 0
        efs.append(mc.StateErrorFunction(pw, rw, tg, weight=0.7, pos_wgt=1.5, rot_wgt=0.8,
                                         rotation_error_type=mc.QUATERNION_LOG_MAP if logmap else mc.ROTATION_MATRIX_DIFFERENCE))
    if "limit" in families:
        efs.append(mc.LimitErrorFunction(weight=0.5, loss_alpha=loss[0], loss_c=loss[1]))
    for fam in ("plane", "halfplane"):  # (after the original families: their random streams stay as they were)
        if fam in families:
            par = rng.integers(0, J, 4).astype(np.int32)
            off = rng.uniform(-1, 1, (4, 3))
            nrm = rng.normal(size=(B, 4, 3)) * rng.uniform(0.5, 2.0, (B, 4, 1))  # not unit length: PlaneDataT normalises
            unit = nrm / np.linalg.norm(nrm, axis=-1, keepdims=True)
            pts = mc.world_points(ch, theta_star, par, off)
            d = np.sum(unit * pts, -1) + 0.2 * rng.normal(size=(B, 4))  # the plane passes near the reachable point, on either side
            w = rng.uniform(0.5, 1.5, 4); w[2] = 0.0
            efs.append(mc.PlaneErrorFunction(par, off, w, np.concatenate([nrm, d[..., None]], -1), above=(fam == "halfplane"), weight=0.8,
                                             loss_alpha=loss[0], loss_c=loss[1]))
    if "model_parameters" in families:
        tw = rng.uniform(0.5, 1.5, n)
        tw[rng.integers(0, n, max(1, n // 4))] = 0.0  # parameters without a target contribute no row
        efs.append(mc.ModelParametersErrorFunction(tw, theta_star + 0.1 * rng.normal(size=(B, n)), weight=0.6))
    theta0 = rng.uniform(-0.3, 0.3, (B, n))
    return ch, efs, theta0, theta_star


def humanoid_problem(B, seed=12347, orientation=True, legacy_weights=True):
    """cfg2 (orientation=False, m=72) / cfg3 (orientation=True, m=126) of BASELINE.json on humanoid72."""
    ch, sets = mc.humanoid72()
    rng = np.random.default_rng(seed)
    n = ch.num_params
    theta_star = np.zeros((B, n))
    theta_star[:, 7:] = rng.uniform(-0.5, 0.5, (B, n - 7))
    theta_star[:, 3:6] = rng.uniform(-0.5, 0.5, (B, 3))
    theta_star[:, 0:3] = rng.uniform(-10.0, 10.0, (B, 3))
    pj = np.array(sets["position_joints"], np.int32)
    poff = rng.uniform(-3.0, 3.0, (len(pj), 3))
    efs = [mc.PositionErrorFunction(pj, poff, np.ones(len(pj)), mc.world_points(ch, theta_star, pj, poff),
                                    weight=mc.PositionErrorFunction.kLegacyWeight if legacy_weights else 1.0)]
    if orientation:
        oj = np.array(sets["orientation_joints"], np.int32)
        ooff = random_unit_quats(rng, (len(oj),))
        efs.append(mc.OrientationErrorFunction(oj, ooff, np.ones(len(oj)), mc.world_rotations(ch, theta_star, oj, ooff),
                                               weight=mc.OrientationErrorFunction.kLegacyWeight if legacy_weights else 1.0))
    return ch, efs, np.zeros((B, n)), theta_star


def bodyhands_problem(B, seed=12349):
    """cfg4: bodyhands300, 200 position (marker) constraints, m = 600, n = 424."""
    ch, sets = mc.bodyhands300()
    rng = np.random.default_rng(seed)
    n = ch.num_params
    theta_star = np.zeros((B, n))
    theta_star[:, 7:] = rng.uniform(-0.4, 0.4, (B, n - 7))
    theta_star[:, 3:6] = rng.uniform(-0.5, 0.5, (B, 3))
    theta_star[:, 0:3] = rng.uniform(-10.0, 10.0, (B, 3))
    pj = np.array(sets["marker_joints"], np.int32)
    poff = rng.uniform(-5.0, 5.0, (len(pj), 3))
    efs = [mc.PositionErrorFunction(pj, poff, np.ones(len(pj)), mc.world_points(ch, theta_star, pj, poff), weight=mc.PositionErrorFunction.kLegacyWeight)]
    return ch, efs, np.zeros((B, n)), theta_star


def chain22_problem(seed=12345):
    """cfg1: createTestCharacter(22), 4 Position constraints on joints {5,10,15,21}."""
    ch = mc.create_test_character(22)
    rng = np.random.default_rng(seed)
    n = ch.num_params
    theta_star = rng.uniform(-0.5, 0.5, (1, n)); theta_star[0, 6] = 0
    pj = np.array([5, 10, 15, 21], np.int32)
    poff = rng.uniform(-1, 1, (4, 3))
    efs = [mc.PositionErrorFunction(pj, poff, np.ones(4), mc.world_points(ch, theta_star, pj, poff), weight=1.0)]
    return ch, efs, np.zeros((1, n)), theta_star


# ---- cfg5: mixed-rig batch -----------------------------------------------------------------------------------------
MIXED_RIGS = (("chain22", 0.25), ("humanoid72", 0.50), ("body150", 0.15), ("bodyhands300", 0.10))


def mixed_rigs():
    """The four rig classes of cfg5 with their canonical marker joints: an instance with c constraints uses the first c of them (so the
    constraint parents of two instances of one rig are prefixes of one another: the bucketing key of the device path)."""
    out = {}
    ch = mc.create_test_character(22)
    out["chain22"] = (ch, [int(j) for j in (np.arange(22) * 5 + 21) % 22])              # every joint once, spread
    ch, sets = mc.humanoid72()
    rest = [j for j in range(1, ch.num_joints) if j not in sets["position_joints"]]
    out["humanoid72"] = (ch, [int(j) for j in list(sets["position_joints"]) + rest])      # end effectors / limb mids first, then the rest
    ch, sets = mc.body150()
    out["body150"] = (ch, list(sets["marker_joints"]))
    ch, sets = mc.bodyhands300()
    out["bodyhands300"] = (ch, list(sets["marker_joints"]))
    return out


def mixed_problem(N, seed=12351, rigs=None):
    """N instances: rig drawn 25/50/15/10 %, constraint count U{4..200} capped by the rig's marker list, offsets and targets per
    instance (reachable poses), theta0 = 0. Returns (rigs dict, list of instance dicts in input order)."""
    rng = np.random.default_rng(seed)
    rigs = rigs or mixed_rigs()
    names = [n for n, _ in MIXED_RIGS]
    probs = np.array([p for _, p in MIXED_RIGS])
    which = rng.choice(len(names), size=N, p=probs)
    counts = rng.integers(4, 201, size=N)
    inst = [None] * N
    for r, name in enumerate(names):  # one batched FK per rig class
        ids = np.nonzero(which == r)[0]
        if ids.size == 0:
            continue
        ch, markers = rigs[name]
        n, cap = ch.num_params, len(markers)
        theta_star = np.zeros((ids.size, n))
        theta_star[:, 7:] = rng.uniform(-0.4, 0.4, (ids.size, n - 7))
        theta_star[:, 3:6] = rng.uniform(-0.5, 0.5, (ids.size, 3))
        theta_star[:, 0:3] = rng.uniform(-10.0, 10.0, (ids.size, 3)) if name != "chain22" else rng.uniform(-1.0, 1.0, (ids.size, 3))
        scale = 1.0 if name == "chain22" else 3.0
        offsets = rng.uniform(-scale, scale, (ids.size, cap, 3))
        targets = mc.world_points(ch, theta_star, np.array(markers, np.int32), offsets)
        for k, i in enumerate(ids):
            c = int(min(counts[i], cap))
            # constraint weights carry PositionErrorFunction::kLegacyWeight (position_error_function.h:64), as in cfg2 / cfg3 / cfg4: with
            # weight 1 and lambda = 0.05 the first Gauss-Newton steps from theta = 0 towards targets ~10 units away overshoot and diverge
            inst[i] = dict(rig=name, parents=np.array(markers[:c], np.int32), offsets=offsets[k, :c].astype(np.float32),
                           weights=np.full(c, mc.PositionErrorFunction.kLegacyWeight if name != "chain22" else 1.0, np.float32),
                           targets=targets[k, :c].astype(np.float32), theta0=np.zeros(n, np.float32), theta_star=theta_star[k])
    return rigs, inst

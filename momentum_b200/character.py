"""Rig and objective descriptions (host side, numpy only).

These plain-data classes mirror the reference types that feed the Gauss-Newton hot path:

* ``Character``  = ``Skeleton`` (momentum/character/skeleton.h:22-77, joint.h:18-76) +
  ``ParameterTransform`` (character/parameter_transform.h:62-184, CSR 7J x n) +
  ``ParameterLimits`` (character/parameter_limits.h:20-138).
* ``PositionErrorFunction`` / ``OrientationErrorFunction`` / ``StateErrorFunction`` /
  ``LimitErrorFunction`` = the constraint data of the same-named reference classes
  (character_solver/position_error_function.h:16-73, orientation_error_function.h:16-108,
  state_error_function.h:35-117, limit_error_function.h:25-119), with a leading batch dimension on
  everything that differs per IK instance (targets, optionally constraint weights).

Synthetic generators follow SURVEY.md §8(d): ``create_test_character`` clones the reference test
fixture (momentum/test/character/character_helpers.cpp:38-55,106-149,213-220); ``humanoid72`` and
``bodyhands300`` are the named benchmark rigs.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

PARAMETERS_PER_JOINT = 7  # character/types.h:21
MAX_MODEL_PARAMETERS = 2048  # ParameterSet = std::bitset<2048>, math/types.h:426-429

# character/parameter_limits.h:20-33
LIMIT_MINMAX, LIMIT_MINMAX_JOINT, LIMIT_MINMAX_JOINT_PASSIVE, LIMIT_LINEAR, LIMIT_LINEAR_JOINT, LIMIT_ELLIPSOID, LIMIT_HALFPLANE = range(7)

# generalized loss special alphas (math/generalized_loss.h:52-60)
LOSS_L2, LOSS_L1, LOSS_CAUCHY = 2.0, 1.0, 0.0
LOSS_WELSCH = float("-inf")

# state_error_function.h:17-32
ROTATION_MATRIX_DIFFERENCE, QUATERNION_LOG_MAP = 0, 1

# error-function kinds (shared numbering with include/momentum_b200.h and oracle/ik_oracle.hpp)
KIND_POSITION, KIND_ORIENTATION, KIND_ORIENTATION_ROTDIFF, KIND_STATE, KIND_LIMIT, KIND_PLANE, KIND_MODEL_PARAMETERS = range(7)


@dataclass
class ParameterLimit:
    """One ``ParameterLimit`` (character/parameter_limits.h:117-127). ``i``/``f`` packing:

    MinMax: i0=parameterIndex, f0,f1=limits. MinMaxJoint: i0=jointIndex, i1=jointParameter, f0,f1.
    Linear: i0=referenceIndex, i1=targetIndex, f0=scale, f1=offset, f2=rangeMin, f3=rangeMax.
    LinearJoint: i0,i1=reference joint/param, i2,i3=target joint/param, f0..f3 as Linear.
    HalfPlane: i0=param1, i1=param2, f0,f1=normal, f2=offset.
    Ellipsoid: i0=ellipsoidParent, i1=parent, f[0:12]=ellipsoid (3x4 row-major), f[12:24]=ellipsoidInv, f[24:27]=offset.
    """

    type: int = LIMIT_MINMAX
    weight: float = 1.0
    i: Sequence[int] = (0, 0, 0, 0)
    f: Sequence[float] = ()

    def packed(self):
        i = np.zeros(4, np.int32)
        i[: len(self.i)] = self.i
        f = np.zeros(27, np.float32)
        f[: len(self.f)] = np.asarray(self.f, np.float32)
        return i, f


@dataclass
class Character:
    parents: np.ndarray  # int32 [J], -1 = root, parents precede children
    offsets: np.ndarray  # float32 [J,3] translationOffset
    prerot: np.ndarray  # float32 [J,4] preRotation (x,y,z,w)
    num_params: int
    pt_outer: np.ndarray  # int32 [7J+1]
    pt_inner: np.ndarray  # int32 [nnz]
    pt_vals: np.ndarray  # float32 [nnz]
    pt_offsets: np.ndarray  # float32 [7J]
    limits: List[ParameterLimit] = field(default_factory=list)
    name: str = "character"

    @property
    def num_joints(self) -> int:
        return int(self.parents.shape[0])

    def validate(self):
        J = self.num_joints
        assert self.offsets.shape == (J, 3) and self.prerot.shape == (J, 4)
        assert self.pt_outer.shape == (7 * J + 1,)
        assert self.num_params <= MAX_MODEL_PARAMETERS
        for j, p in enumerate(self.parents):
            assert -1 <= p < j, "skeleton must be topologically sorted (skeleton.h:23-24)"
        assert self.pt_inner.size == 0 or self.pt_inner.max() < self.num_params

    def depth(self) -> np.ndarray:
        d = np.zeros(self.num_joints, np.int32)
        for j, p in enumerate(self.parents):
            d[j] = 0 if p < 0 else d[p] + 1
        return d


def _csr_from_triplets(rows, n_params, triplets):
    trip = sorted(triplets, key=lambda t: (t[0], t[1]))
    outer = np.zeros(rows + 1, np.int32)
    for r, _, _ in trip:
        outer[r + 1] += 1
    outer = np.cumsum(outer).astype(np.int32)
    inner = np.array([t[1] for t in trip], np.int32)
    vals = np.array([t[2] for t in trip], np.float32)
    return outer, inner, vals


def create_test_character(num_joints: int = 3) -> Character:
    """Clone of ``createTestCharacter`` (momentum/test/character/character_helpers.cpp:224-244):
    Y-axis chain, unit offsets, identity pre-rotations, n = 9 + (J-2) model parameters
    {root tx,ty,tz,rx,ry,rz, scale_global, joint1_rx, shared_rz(0.5*j1.rz + 0.5*j2.rz), jointK_rx}
    and one MinMax limit on parameter 0 (:213-220)."""
    assert num_joints >= 3
    J = num_joints
    parents = np.arange(-1, J - 1, dtype=np.int32)
    offsets = np.zeros((J, 3), np.float32)
    offsets[1:, 1] = 1.0
    prerot = np.zeros((J, 4), np.float32)
    prerot[:, 3] = 1.0
    trip = [(k, k, 1.0) for k in range(7)]
    trip.append((1 * 7 + 3, 7, 1.0))
    trip.append((1 * 7 + 5, 8, 0.5))
    trip.append((2 * 7 + 5, 8, 0.5))
    for j in range(2, J):
        trip.append((j * 7 + 3, 9 + j - 2, 1.0))
    n = 9 + J - 2
    outer, inner, vals = _csr_from_triplets(7 * J, n, trip)
    limits = [ParameterLimit(LIMIT_MINMAX, 1.0, (0,), (-0.1, 0.1))]
    return Character(parents, offsets, prerot, n, outer, inner, vals, np.zeros(7 * J, np.float32), limits, f"chain{J}")


def _random_prerot(rng, max_angle_deg=30.0):
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(0, max_angle_deg))
    s = np.sin(ang / 2)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, np.cos(ang / 2)], np.float32)


class _TreeBuilder:
    def __init__(self, rng):
        self.rng = rng
        self.parents, self.offsets, self.prerot, self.names = [], [], [], []

    def add(self, name, parent, length, direction=None):
        rng = self.rng
        if direction is None:
            direction = rng.normal(size=3)
        direction = np.asarray(direction, np.float64)
        direction = direction / np.linalg.norm(direction)
        self.parents.append(parent)
        self.offsets.append((direction * length).astype(np.float32))
        self.prerot.append(_random_prerot(rng) if parent >= 0 else np.array([0, 0, 0, 1], np.float32))
        self.names.append(name)
        return len(self.parents) - 1

    def chain(self, prefix, parent, lengths, direction):
        ids = []
        for k, L in enumerate(lengths):
            d = np.asarray(direction, np.float64) + 0.15 * self.rng.normal(size=3)
            parent = self.add(f"{prefix}{k}", parent, L, d)
            ids.append(parent)
        return ids


def humanoid72(seed: int = 12346) -> "tuple[Character, dict]":
    """72-joint humanoid, n = 220: pelvis root (6 DOF + global scale) and
    rx,ry,rz on each of the other 71 joints. Lengths are in centimetre-like units so that the
    reference's legacy weights (Position 1e-4, Orientation 1e-1) give O(1) normal-equation entries
    against the default damping 0.05. Returns (character, named joint sets)."""
    rng = np.random.default_rng(seed)
    tb = _TreeBuilder(rng)
    U = lambda a, b: float(rng.uniform(a, b))
    root = tb.add("pelvis", -1, 0.0, (0, 1, 0))
    tb.offsets[0][:] = 0
    spine = tb.chain("spine", root, [U(8, 14) for _ in range(4)], (0, 1, 0))
    neck = tb.add("neck", spine[-1], U(8, 12), (0, 1, 0))
    head = tb.add("head", neck, U(8, 12), (0, 1, 0))
    sets = {"pelvis": root, "head": head, "neck": neck}
    fingers_tips = []
    for side, sx in (("l", 1.0), ("r", -1.0)):
        clav = tb.add(f"{side}_clavicle", spine[-1], U(12, 18), (sx, 0.2, 0))
        sho = tb.add(f"{side}_shoulder", clav, U(5, 8), (sx, 0, 0))
        elb = tb.add(f"{side}_elbow", sho, U(25, 32), (sx, -0.2, 0))
        twist = tb.add(f"{side}_forearm_twist", elb, U(11, 14), (sx, 0, 0))
        wrist = tb.add(f"{side}_wrist", twist, U(11, 14), (sx, 0, 0))
        for fi in range(5):
            ids = tb.chain(f"{side}_finger{fi}_", wrist, [U(6, 9)] + [U(2, 4) for _ in range(3)], (sx, 0.1 * (fi - 2), 0.3 * (fi - 2)))
            fingers_tips.append(ids[-1])
        sets[f"{side}_shoulder"], sets[f"{side}_elbow"], sets[f"{side}_wrist"] = sho, elb, wrist
    for side, sx in (("l", 1.0), ("r", -1.0)):
        hip = tb.add(f"{side}_hip", root, U(9, 12), (sx, -0.3, 0))
        knee = tb.add(f"{side}_knee", hip, U(38, 46), (0, -1, 0))
        ankle = tb.add(f"{side}_ankle", knee, U(36, 44), (0, -1, 0))
        ball = tb.add(f"{side}_ball", ankle, U(10, 14), (0, -0.3, 1))
        toe = tb.add(f"{side}_toe", ball, U(5, 8), (0, 0, 1))
        sets[f"{side}_knee"], sets[f"{side}_ankle"], sets[f"{side}_toe"] = knee, ankle, toe
    helpers = [spine[1], spine[2], sets["l_elbow"], sets["r_elbow"], head]
    for k, h in enumerate(helpers):
        tb.add(f"helper{k}", h, U(4, 8))
    J = len(tb.parents)
    assert J == 72, J
    trip = [(k, k, 1.0) for k in range(7)]
    col = 7
    for j in range(1, J):
        for d in range(3):
            trip.append((j * 7 + 3 + d, col, 1.0))
            col += 1
    n = col
    assert n == 220
    outer, inner, vals = _csr_from_triplets(7 * J, n, trip)
    ch = Character(np.array(tb.parents, np.int32), np.stack(tb.offsets).astype(np.float32), np.stack(tb.prerot).astype(np.float32), n, outer, inner, vals, np.zeros(7 * J, np.float32), [], "humanoid72")
    sets["fingertips"] = fingers_tips
    pos_joints = [head, neck, sets["l_wrist"], sets["r_wrist"], sets["l_elbow"], sets["r_elbow"], sets["l_shoulder"], sets["r_shoulder"],
                  sets["l_ankle"], sets["r_ankle"], sets["l_knee"], sets["r_knee"], sets["l_toe"], sets["r_toe"]] + fingers_tips
    assert len(pos_joints) == 24
    sets["position_joints"] = pos_joints
    sets["orientation_joints"] = [root, head, sets["l_wrist"], sets["r_wrist"], sets["l_ankle"], sets["r_ankle"]]
    ch.validate()
    return ch, sets


def bodyhands_rig(num_chains: int = 60, chain_len: int = 4, seed: int = 12348, name: str = "bodyhands300") -> "tuple[Character, dict]":
    """Body + helper-chain rig: 60 body joints (root 6 DOF + scale, 3 rotation DOF on the other 59) + ``num_chains`` chains of
    ``chain_len`` one-DOF (rx) finger/helper joints; every 8th helper joint's rx row is additionally driven by its parent's parameter
    with weight 0.5 (mirrors ``shared_rz`` of the reference fixture, character_helpers.cpp:137-138).
    60 x 4 = the 300-joint body+hands rig with n = 424 of SURVEY.md §8d; 30 x 3 = the 150-joint rig of cfg5 (n = 274)."""
    rng = np.random.default_rng(seed)
    tb = _TreeBuilder(rng)
    U = lambda a, b: float(rng.uniform(a, b))
    root = tb.add("pelvis", -1, 0.0, (0, 1, 0))
    tb.offsets[0][:] = 0
    spine = tb.chain("spine", root, [U(7, 11) for _ in range(5)], (0, 1, 0))
    neck = tb.chain("neck", spine[-1], [U(5, 7), U(5, 7)], (0, 1, 0))
    head = tb.add("head", neck[-1], U(8, 12), (0, 1, 0))
    attach = []
    for side, sx in (("l", 1.0), ("r", -1.0)):
        arm = tb.chain(f"{side}_arm", spine[-1], [U(12, 18), U(5, 8), U(13, 16), U(13, 16), U(11, 14), U(11, 14)], (sx, 0, 0))
        leg = tb.chain(f"{side}_leg", root, [U(9, 12), U(19, 23), U(19, 23), U(18, 22), U(18, 22), U(10, 14), U(5, 8)], (0.2 * sx, -1, 0))
        attach += [arm[-1]] * 5 + [leg[-1]] * 3 + [leg[-2]] * 2
    body_leaf_parents = spine + neck + [head]
    k = 0
    while len(tb.parents) < 60:
        tb.add(f"body_helper{k}", body_leaf_parents[k % len(body_leaf_parents)], U(4, 9))
        k += 1
    n_body = len(tb.parents)
    assert n_body == 60
    attach += [head] * 4 + spine * 4
    body_all = list(range(1, n_body))
    k = 0
    while len(attach) < 60:  # remaining helper chains hang off body joints round-robin
        attach.append(body_all[(7 * k) % len(body_all)])
        k += 1
    chain_parents = attach[:num_chains]
    assert len(chain_parents) == num_chains
    helper_ids = []
    for ci, par in enumerate(chain_parents):
        ids = tb.chain(f"h{ci}_", par, [U(2, 6) for _ in range(chain_len)], rng.normal(size=3))
        helper_ids += ids
    J = len(tb.parents)
    assert J == 60 + num_chains * chain_len, J
    trip = [(k, k, 1.0) for k in range(7)]
    col = 7
    for j in range(1, n_body):
        for d in range(3):
            trip.append((j * 7 + 3 + d, col, 1.0))
            col += 1
    own = {}
    for j in helper_ids:
        own[j] = col
        trip.append((j * 7 + 3, col, 1.0))
        col += 1
    for idx, j in enumerate(helper_ids):
        p = tb.parents[j]
        if idx % 8 == 7 and p in own:
            trip.append((j * 7 + 3, own[p], 0.5))
    n = col
    assert n == 7 + 3 * 59 + num_chains * chain_len, n
    outer, inner, vals = _csr_from_triplets(7 * J, n, trip)
    ch = Character(np.array(tb.parents, np.int32), np.stack(tb.offsets).astype(np.float32), np.stack(tb.prerot).astype(np.float32), n, outer, inner, vals, np.zeros(7 * J, np.float32), [], name)
    ch.validate()
    sets = {"marker_joints": [int(j) for j in np.round(np.linspace(1, J - 1, min(200, J - 1))).astype(int)]}
    return ch, sets


def bodyhands300(seed: int = 12348) -> "tuple[Character, dict]":
    """300-joint body+hands rig, n = 424."""
    return bodyhands_rig(60, 4, seed, "bodyhands300")


def body150(seed: int = 12350) -> "tuple[Character, dict]":
    """150-joint rig of the cfg5 mix (60 body joints + 30 helper chains of three), n = 274."""
    return bodyhands_rig(30, 3, seed, "body150")


# ------------------------------------------------------------------------------------------------
# Error-function data (batched)
# ------------------------------------------------------------------------------------------------
@dataclass
class PositionErrorFunction:
    """PositionErrorFunctionT constraints (position_error_function.h:16-29): per constraint parent
    joint, offset in the parent frame, weight; ``targets`` carries the batch dimension [B, nc, 3]."""

    parents: np.ndarray
    offsets: np.ndarray  # [nc,3]
    weights: np.ndarray  # [nc] (shared) — ConstraintData::weight is float
    targets: np.ndarray  # [B,nc,3]
    weight: float = 1.0  # SkeletonErrorFunctionT::weight_
    loss_alpha: float = LOSS_L2
    loss_c: float = 1.0
    kind: int = KIND_POSITION
    instance_offsets: Optional[np.ndarray] = None  # [B,nc,3]: offsets per batch element (tensor_ik.cpp:136-140 builds them per element)
    kLegacyWeight = 1e-4  # position_error_function.h:64


@dataclass
class OrientationErrorFunction:
    """OrientationErrorFunctionT / OrientationRotDiffErrorFunctionT (orientation_error_function.h:16-108);
    quaternions are (x,y,z,w) and are normalised on entry as in OrientationDataT's constructor."""

    parents: np.ndarray
    offsets: np.ndarray  # [nc,4]
    weights: np.ndarray  # [nc]
    targets: np.ndarray  # [B,nc,4]
    weight: float = 1.0
    loss_alpha: float = LOSS_L2
    loss_c: float = 1.0
    rot_diff: bool = False
    kLegacyWeight = 1e-1  # orientation_error_function.h:63

    @property
    def kind(self):
        return KIND_ORIENTATION_ROTDIFF if self.rot_diff else KIND_ORIENTATION


@dataclass
class StateErrorFunction:
    """StateErrorFunctionT (state_error_function.h:35-117): per-joint position/rotation target
    weights, global pos/rot weights; ``targets`` [B, J, 8] = (t, q xyzw, s) per joint."""

    pos_weights: np.ndarray  # [J]
    rot_weights: np.ndarray  # [J]
    targets: np.ndarray  # [B,J,8]
    weight: float = 1.0
    pos_wgt: float = 1.0
    rot_wgt: float = 1.0
    rotation_error_type: int = ROTATION_MATRIX_DIFFERENCE
    kind: int = KIND_STATE


@dataclass
class LimitErrorFunction:
    """LimitErrorFunctionT over the character's ParameterLimits (limit_error_function.h:25-119)."""

    weight: float = 1.0
    loss_alpha: float = LOSS_L2
    loss_c: float = 1.0
    kind: int = KIND_LIMIT


@dataclass
class PlaneErrorFunction:
    """PlaneErrorFunctionT (plane_error_function.h:20-101, .cpp:49-70): signed distance of T_parent * offset to the plane
    (normal, d), one row per constraint; ``above`` = half-plane mode (only penetration is penalised). ``targets`` [B, nc, 4] =
    (normal xyz, d) per instance; normals are normalised on entry as in PlaneDataT's constructor."""

    parents: np.ndarray
    offsets: np.ndarray  # [nc,3]
    weights: np.ndarray  # [nc]
    targets: np.ndarray  # [B,nc,4]
    above: bool = False
    weight: float = 1.0
    loss_alpha: float = LOSS_L2
    loss_c: float = 1.0
    kind: int = KIND_PLANE
    kLegacyWeight = 1e-4  # plane_error_function.h:83


@dataclass
class ModelParametersErrorFunction:
    """ModelParametersErrorFunctionT (model_parameters_error_function.h/.cpp): w_i (theta_i - target_i) on every enabled parameter
    with target weight > 0, scaled by weight * kMotionWeight (1e-1). ``targets`` [B, n] per instance, ``target_weights`` [n] shared."""

    target_weights: np.ndarray  # [n]
    targets: np.ndarray  # [B,n]
    weight: float = 1.0
    kind: int = KIND_MODEL_PARAMETERS
    kMotionWeight = 1e-1  # model_parameters_error_function.h:61


def jacobian_size(character: Character, ef) -> int:
    """getJacobianSize() of each family (joint_error_function-inl.h:300-302,
    state_error_function.cpp:394-404, limit_error_function.cpp:1138-1161)."""
    if ef.kind == KIND_POSITION:
        return 3 * len(ef.parents)
    if ef.kind in (KIND_ORIENTATION, KIND_ORIENTATION_ROTDIFF):
        return 9 * len(ef.parents)
    if ef.kind == KIND_STATE:
        active = int(np.count_nonzero((np.asarray(ef.pos_weights) != 0) | (np.asarray(ef.rot_weights) != 0)))
        return active * (6 if ef.rotation_error_type == QUATERNION_LOG_MAP else 12)
    if ef.kind == KIND_PLANE:
        return len(ef.parents)
    if ef.kind == KIND_MODEL_PARAMETERS:  # model_parameters_error_function.cpp:93-95
        return int(np.count_nonzero(np.asarray(ef.target_weights) > 0))
    if ef.kind == KIND_LIMIT:
        return sum(0 if l.type == LIMIT_MINMAX_JOINT_PASSIVE else (3 if l.type == LIMIT_ELLIPSOID else 1) for l in character.limits)
    raise ValueError(ef.kind)


# ------------------------------------------------------------------------------------------------
# numpy forward kinematics (float64) — used only to synthesise reachable targets for benchmarks/tests
# ------------------------------------------------------------------------------------------------
def _qmul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], -1)


def _qrot(q, v):
    u = q[..., :3]
    uv = np.cross(u, v)
    uv = uv + uv
    return v + q[..., 3:4] * uv + np.cross(u, uv)


def forward_kinematics(ch: Character, theta: np.ndarray):
    """Batched FK in float64 following joint_state.cpp:22-65 / transform.h:124-129.
    theta [B,n] -> (t [B,J,3], q [B,J,4], s [B,J])."""
    theta = np.asarray(theta, np.float64)
    B = theta.shape[0]
    J = ch.num_joints
    jp = np.zeros((B, 7 * J))
    rows = np.repeat(np.arange(7 * J), np.diff(ch.pt_outer))
    np.add.at(jp, (slice(None), rows), theta[:, ch.pt_inner] * ch.pt_vals.astype(np.float64))
    jp += ch.pt_offsets.astype(np.float64)
    jp = jp.reshape(B, J, 7)
    t = np.zeros((B, J, 3)); q = np.zeros((B, J, 4)); s = np.zeros((B, J))
    for j in range(J):
        p = jp[:, j]
        ql = np.broadcast_to(ch.prerot[j].astype(np.float64), (B, 4)).copy()
        for k in (2, 1, 0):
            r = np.zeros((B, 4)); r[:, k] = np.sin(0.5 * p[:, 3 + k]); r[:, 3] = np.cos(0.5 * p[:, 3 + k])
            ql = _qmul(ql, r)
        tl = ch.offsets[j].astype(np.float64) + p[:, :3]
        sl = np.exp2(p[:, 6])
        par = ch.parents[j]
        if par < 0:
            t[:, j], q[:, j], s[:, j] = tl, ql, sl
        else:
            t[:, j] = t[:, par] + _qrot(q[:, par], s[:, par, None] * tl)
            q[:, j] = _qmul(q[:, par], ql)
            s[:, j] = s[:, par] * sl
    return t, q, s


def world_points(ch: Character, theta, parents, offsets):
    """offsets [nc,3] (shared by the batch) or [B,nc,3] (per instance)."""
    t, q, s = forward_kinematics(ch, theta)
    parents = np.asarray(parents)
    off = np.asarray(offsets, np.float64)
    if off.ndim == 2:
        off = off[None]
    return t[:, parents] + _qrot(q[:, parents], s[:, parents, None] * off)


def world_rotations(ch: Character, theta, parents, offsets_q):
    _, q, _ = forward_kinematics(ch, theta)
    parents = np.asarray(parents)
    return _qmul(q[:, parents], np.broadcast_to(np.asarray(offsets_q, np.float64)[None], (q.shape[0], len(parents), 4)))

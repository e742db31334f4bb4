// Per-lane building blocks of the FK / residual / Jacobian sweep. Every function is
// __host__ __device__ so the same code is exercised by the CPU lane-emulation test
// (tests/emu) before it runs in the sm_100a kernels of ik_kernels.cu.
//
// Reference math (file:line relative to /root/reference/momentum):
//   character/joint_state.cpp:22-82, character/skeleton_state.cpp:87-121,
//   character/parameter_transform.cpp:110-123, math/transform.h:124-129,165-167,193-195,
//   character_solver/joint_error_function-inl.h:179-297, position_/orientation_/state_/limit_error_function.cpp,
//   math/generalized_loss.cpp:24-160, math/utility.cpp:72-180.
#pragma once

#include <cmath>
#include <cfloat>

#include "ik_types.h"

namespace mb2 {

struct F3 {
  float x, y, z;
};
struct Q4 {
  float x, y, z, w;
};

MB2_HD F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
MB2_HD F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
MB2_HD F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
MB2_HD F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
MB2_HD F3 operator*(float s, F3 a) { return f3(a.x * s, a.y * s, a.z * s); }
MB2_HD float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MB2_HD F3 cross(F3 a, F3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
MB2_HD float comp(F3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

MB2_HD Q4 q4(float x, float y, float z, float w) { Q4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
// Eigen quaternion product
MB2_HD Q4 qmul(Q4 a, Q4 b) {
  return q4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
// Eigen QuaternionBase::_transformVector
MB2_HD F3 qrot(Q4 q, F3 v) {
  F3 u = f3(q.x, q.y, q.z);
  F3 uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}
MB2_HD Q4 qconj(Q4 q) { return q4(-q.x, -q.y, -q.z, q.w); }
MB2_HD Q4 qnormalized(Q4 q) {
  const float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return q4(q.x / n, q.y / n, q.z / n, q.w / n);
}
// Eigen QuaternionBase::toRotationMatrix; m[3*col + row]
MB2_HD void qmat(Q4 q, float* m) {
  const float tx = 2.f * q.x, ty = 2.f * q.y, tz = 2.f * q.z;
  const float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  m[0] = 1.f - (tyy + tzz); m[3] = txy - twz; m[6] = txz + twy;
  m[1] = txy + twz; m[4] = 1.f - (txx + tzz); m[7] = tyz - twx;
  m[2] = txz - twy; m[5] = tyz + twx; m[8] = 1.f - (txx + tyy);
}
MB2_HD F3 qmatcol(Q4 q, int c) { // column c of toRotationMatrix(q)
  float m[9];
  qmat(q, m);
  return f3(m[3 * c], m[3 * c + 1], m[3 * c + 2]);
}

MB2_HD F3 ld3(const float* p) { return f3(p[0], p[1], p[2]); }
MB2_HD Q4 ld4(const float* p) { return q4(p[0], p[1], p[2], p[3]); }

// ---- GeneralizedLossT (math/generalized_loss.cpp:104-155) ----
MB2_HD float lossValue(const EfDesc& e, float s) {
  switch (e.lossType) {
    case kLossL2: return s * e.invC2;
    case kLossL1: return sqrtf(s * e.invC2 + 1.f) - 1.f;
    case kLossCauchy: return logf(0.5f * (s * e.invC2) + 1.f);
    case kLossWelsch: return 1.f - expf(-0.5f * (s * e.invC2));
    default: return (powf(s * e.invC2 / fabsf(e.alpha - 2.f) + 1.f, 0.5f * e.alpha) - 1.f) * fabsf(e.alpha - 2.f) / e.alpha;
  }
}
MB2_HD float lossDeriv(const EfDesc& e, float s) {
  switch (e.lossType) {
    case kLossL2: return e.invC2;
    case kLossL1: return 0.5f * e.invC2 / sqrtf(s * e.invC2 + 1.f);
    case kLossCauchy: return e.invC2 / (e.invC2 * s + 2.f);
    case kLossWelsch: return 0.5f * e.invC2 * expf(-0.5f * (s * e.invC2));
    default: return 0.5f * e.invC2 * powf(s * e.invC2 / fabsf(e.alpha - 2.f) + 1.f, 0.5f * e.alpha - 1.f);
  }
}

// ---- ParameterTransformT::apply, one row (parameter_transform.cpp:122) ----
MB2_HD float jointParameterRow(const FunctionTables& T, int row, const float* theta) {
  float s = 0.f;
  for (int k = T.ptOuter[row]; k < T.ptOuter[row + 1]; ++k) s += T.ptVals[k] * theta[T.ptInner[k]];
  return s + T.ptOffsets[row];
}

// ---- JointStateT::set for joint j (joint_state.cpp:22-65). js: kJointStateStride floats per joint ----
template <bool kDeriv>
MB2_HD void fkJoint(const FunctionTables& T, int j, const float* jp, float* js) {
  const float* p = jp + j * kParametersPerJoint;
  const int par = T.parent[j];
  F3 tp = f3(0.f, 0.f, 0.f);
  Q4 qp = q4(0.f, 0.f, 0.f, 1.f);
  float sp = 1.f;
  if (par >= 0) {
    const float* ps = js + par * kJointStateStride;
    tp = ld3(ps); qp = ld4(ps + 3); sp = ps[7];
  }
  Q4 ql = ld4(T.prerot + 4 * j);
  float* out = js + j * kJointStateStride;
  for (int index = 2; index >= 0; --index) { // :51-58
    if (kDeriv) {
      const F3 axis = f3(index == 0 ? 1.f : 0.f, index == 1 ? 1.f : 0.f, index == 2 ? 1.f : 0.f);
      const F3 a = qrot(qmul(qp, ql), axis);
      out[8 + 3 * index] = a.x; out[9 + 3 * index] = a.y; out[10 + 3 * index] = a.z;
    }
    const float ha = 0.5f * p[3 + index];
    float sn, cs;
#if defined(__CUDA_ARCH__)
    sincosf(ha, &sn, &cs);
#else
    sn = sinf(ha); cs = cosf(ha);
#endif
    ql = qmul(ql, q4(index == 0 ? sn : 0.f, index == 1 ? sn : 0.f, index == 2 ? sn : 0.f, cs));
  }
  const F3 tl = f3(T.offset[3 * j] + p[0], T.offset[3 * j + 1] + p[1], T.offset[3 * j + 2] + p[2]); // :44
  const float sl = exp2f(p[6]); // :62
  const F3 t = tp + qrot(qp, sp * tl); // transform.h:125
  const Q4 q = qmul(qp, ql);
  out[0] = t.x; out[1] = t.y; out[2] = t.z;
  out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
  out[7] = sp * sl;
}

// ---- The same joint state in three data-parallel passes (what the kernels run) -----------------------------------------------
// JointStateT::set mixes work that needs the parent (two compositions) with work that does not (three sin/cos pairs, the local
// rotation chain, exp2). Splitting it lets every joint of the skeleton do the expensive part at once, leaves ~50 dependent
// flops per tree level, and turns the derivative axes into one more flat pass:
//   fkLocal   (all joints)         t_l, q_l = preRot Rz Ry Rx, s_l, and the DOF axes in the PARENT frame: a_2 = preRot z, a_1 = (preRot Rz) y,
//                                  a_0 = (preRot Rz Ry) x                                                   (joint_state.cpp:44-62)
//   fkCompose (level by level)     t = t_p + q_p (s_p t_l), q = q_p q_l, s = s_p s_l                          (transform.h:124-129)
//   fkAxis    (all joints x 3)     rotationAxis.col(i) = q_p a_i   [= (q_p preRot ...) e_i of joint_state.cpp:51-55: one rotation of a
//                                  rotated vector instead of a rotation by a quaternion product; equal up to rounding]
// fkJoint above stays the statement-by-statement form (CPU emulation of the oracle order, tests).
template <bool kDeriv>
MB2_HD void fkLocalFromParameters(const FunctionTables& T, int j, const float* p, float* js);
template <bool kDeriv>
MB2_HD void fkLocal(const FunctionTables& T, int j, const float* jp, float* js) { fkLocalFromParameters<kDeriv>(T, j, jp + j * kParametersPerJoint, js); }
// the joint's seven parameters straight from theta (ParameterTransform rows 7 j .. 7 j + 6): no [7 J] array in shared memory, and the
// transform is spread over lanes = joints (three rounds for 72 joints) instead of lanes = rows (sixteen rounds of dependent loads)
template <bool kDeriv>
MB2_HD void fkLocalFromTheta(const FunctionTables& T, int j, const float* theta, float* js) {
  float p[kParametersPerJoint];
#pragma unroll
  for (int r = 0; r < kParametersPerJoint; ++r) p[r] = jointParameterRow(T, j * kParametersPerJoint + r, theta);
  fkLocalFromParameters<kDeriv>(T, j, p, js);
}
// jp == nullptr: the caller keeps no joint-parameter array (same value, recomputed from theta)
MB2_HD float jointParameterAt(const FunctionTables& T, const float* jp, const float* theta, int row) { return jp != nullptr ? jp[row] : jointParameterRow(T, row, theta); }
template <bool kDeriv>
MB2_HD void fkLocalFromParameters(const FunctionTables& T, int j, const float* p, float* js) {
  Q4 ql = ld4(T.prerot + 4 * j);
  float* out = js + j * kJointStateStride;
#pragma unroll
  for (int index = 2; index >= 0; --index) {
    if (kDeriv) {
      const F3 a = qrot(ql, f3(index == 0 ? 1.f : 0.f, index == 1 ? 1.f : 0.f, index == 2 ? 1.f : 0.f));
      out[8 + 3 * index] = a.x; out[9 + 3 * index] = a.y; out[10 + 3 * index] = a.z;
    }
    const float ha = 0.5f * p[3 + index];
    float sn, cs;
#if defined(__CUDA_ARCH__)
    sincosf(ha, &sn, &cs);
#else
    sn = sinf(ha); cs = cosf(ha);
#endif
    ql = qmul(ql, q4(index == 0 ? sn : 0.f, index == 1 ? sn : 0.f, index == 2 ? sn : 0.f, cs));
  }
  out[0] = T.offset[3 * j] + p[0]; out[1] = T.offset[3 * j + 1] + p[1]; out[2] = T.offset[3 * j + 2] + p[2];
  out[3] = ql.x; out[4] = ql.y; out[5] = ql.z; out[6] = ql.w;
  out[7] = exp2f(p[6]);
}
// the parent (if any) already holds its world transform
MB2_HD void fkCompose(const FunctionTables& T, int j, float* js) {
  const int par = T.parent[j];
  if (par < 0) return; // identity parent: world = local, bit for bit
  const float* ps = js + par * kJointStateStride;
  float* out = js + j * kJointStateStride;
  const F3 tp = ld3(ps);
  const Q4 qp = ld4(ps + 3);
  const float sp = ps[7];
  const F3 t = tp + qrot(qp, sp * ld3(out));
  const Q4 q = qmul(qp, ld4(out + 3));
  out[0] = t.x; out[1] = t.y; out[2] = t.z;
  out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
  out[7] = sp * out[7];
}
// after every joint is composed: world direction of DOF axis `index` of joint j
MB2_HD void fkAxis(const FunctionTables& T, int j, int index, float* js) {
  const int par = T.parent[j];
  if (par < 0) return;
  float* a = js + j * kJointStateStride + 8 + 3 * index;
  const F3 w = qrot(ld4(js + par * kJointStateStride + 3), ld3(a));
  a[0] = w.x; a[1] = w.y; a[2] = w.z;
}

// translationAxis of joint a = parent.toLinear() (joint_state.cpp:36-42), column d
MB2_HD F3 translationAxisCol(const FunctionTables& T, const float* js, int a, int d) {
  const int par = T.parent[a];
  if (par < 0) return f3(d == 0 ? 1.f : 0.f, d == 1 ? 1.f : 0.f, d == 2 ? 1.f : 0.f);
  const float* ps = js + par * kJointStateStride;
  return qmatcol(ld4(ps + 3), d) * ps[7];
}
MB2_HD F3 rotationAxisCol(const float* js, int a, int d) { return ld3(js + a * kJointStateStride + 8 + 3 * d); }

// derivative of a world point v attached below joint a w.r.t. joint-parameter dof d
// (joint_state.cpp:68-82 + joint_error_function-inl.h:240-291)
MB2_HD F3 pointDerivative(const FunctionTables& T, const float* js, int a, int d, F3 v) {
  if (d < 3) return translationAxisCol(T, js, a, d);
  const F3 off = v - ld3(js + a * kJointStateStride);
  if (d < 6) return cross(rotationAxisCol(js, a, d - 3), off);
  return off * kLn2;
}

// ---- quaternion log map (math/utility.cpp:72-180) ----
MB2_HD F3 quaternionLogMap(Q4 q) {
  const Q4 qn = qnormalized(q);
  const F3 vec = f3(qn.x, qn.y, qn.z);
  const float vn = sqrtf(dot(vec, vec));
  if (vn < 3.5e-4f) {
    if (qn.w > 0.f) return vec * (2.f * (1.f + dot(vec, vec) / 6.f));
    return f3(kPi, 0.f, 0.f);
  }
  const float theta = 2.f * atan2f(vn, qn.w);
  return vec * (theta / vn);
}
MB2_HD void quaternionLogMapDerivative(Q4 q, float* jac /* [row*4 + col], cols x,y,z,w */) {
  const Q4 qn = qnormalized(q);
  const float v[3] = {qn.x, qn.y, qn.z};
  const float w = qn.w;
  const float vn2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const float vn = sqrtf(vn2);
  if (vn < 3.5e-4f) {
    const float scale = 2.f * (1.f + vn2 / 6.f);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) jac[4 * i + j] = (i == j ? scale : 0.f) + 2.f * v[i] * v[j] / 3.f;
      jac[4 * i + 3] = 0.f;
    }
    return;
  }
  const float theta = 2.f * atan2f(vn, w);
  const float scale = theta / vn;
  const float denom = w * w + vn * vn;
  const float dthetaDw = -2.f * vn / denom;
  for (int j = 0; j < 3; ++j) {
    const float dthetaDvj = 2.f * w * v[j] / (vn * denom);
    const float dScale = dthetaDvj / vn - theta * v[j] / (vn * vn * vn);
    for (int i = 0; i < 3; ++i) jac[4 * i + j] = dScale * v[i] + (i == j ? scale : 0.f);
  }
  for (int i = 0; i < 3; ++i) jac[4 * i + 3] = dthetaDw / vn * v[i];
}
// state_error_function.cpp:33-66
MB2_HD F3 logMapRelativeDerivativeQ1(Q4 q1, Q4 q2, F3 dir, const float* dLogDq) {
  const F3 vHalf = dir * 0.5f;
  const F3 q1v = f3(q1.x, q1.y, q1.z);
  const float dq1w = -dot(vHalf, q1v);
  const F3 dq1v = vHalf * q1.w + cross(vHalf, q1v);
  const F3 mq2v = f3(-q2.x, -q2.y, -q2.z);
  const float dqRelW = q2.w * dq1w - dot(mq2v, dq1v);
  const F3 dqRelV = q2.w * dq1v + dq1w * mq2v + cross(mq2v, dq1v);
  return f3(dLogDq[0] * dqRelV.x + dLogDq[1] * dqRelV.y + dLogDq[2] * dqRelV.z + dLogDq[3] * dqRelW,
            dLogDq[4] * dqRelV.x + dLogDq[5] * dqRelV.y + dLogDq[6] * dqRelV.z + dLogDq[7] * dqRelW,
            dLogDq[8] * dqRelV.x + dLogDq[9] * dqRelV.y + dLogDq[10] * dqRelV.z + dLogDq[11] * dqRelW);
}

MB2_HD bool limitInRange(float rangeMin, float rangeMax, float v) { // parameter_limits.cpp:105-123
  if (rangeMin == 0.f && rangeMax == 0.f) return true;
  return v >= rangeMin && v < rangeMax;
}

// Ellipsoid limit evaluation (limit_error_function.cpp:713-722): returns position, diff
MB2_HD void evalEllipsoid(const FunctionTables& T, const UnitDesc& u, const float* js, F3& position, F3& diff) {
  const float* d = T.limitData + u.extra;
  const float* ps = js + u.joint * kJointStateStride;
  position = ld3(ps) + qrot(ld4(ps + 3), ps[7] * f3(d[24], d[25], d[26]));
  const float* es = js + u.i[0] * kJointStateStride; // ellipsoidParent
  const F3 et = ld3(es);
  const Q4 eq = ld4(es + 3);
  const float esc = es[7];
  // TransformT::inverse (math/transform.cpp:93-101), Eigen quaternion inverse = conj / squaredNorm
  const float n2 = eq.x * eq.x + eq.y * eq.y + eq.z * eq.z + eq.w * eq.w;
  const Q4 iq = q4(-eq.x / n2, -eq.y / n2, -eq.z / n2, eq.w / n2);
  const float is = 1.f / esc;
  const F3 it = (qrot(iq, et) * is) * -1.f;
  const F3 local = it + qrot(iq, is * position);
  F3 ep = f3(d[12] * local.x + d[13] * local.y + d[14] * local.z + d[15], d[16] * local.x + d[17] * local.y + d[18] * local.z + d[19],
             d[20] * local.x + d[21] * local.y + d[22] * local.z + d[23]);
  const float inv = 1.f / sqrtf(dot(ep, ep));
  ep = ep * inv;
  const F3 proj = f3(d[0] * ep.x + d[1] * ep.y + d[2] * ep.z + d[3], d[4] * ep.x + d[5] * ep.y + d[6] * ep.z + d[7],
                     d[8] * ep.x + d[9] * ep.y + d[10] * ep.z + d[11]);
  diff = position - (et + qrot(eq, esc * proj));
}

constexpr float kLimitWeight = 10.f;        // limit_error_function.h:91
constexpr float kLimitPositionWeight = 1e-4f; // limit_error_function.cpp:21
constexpr float kStatePositionWeight = 1e-3f; // state_error_function.h:115
constexpr float kStateOrientationWeight = 1.f; // state_error_function.h:116

// ------------------------------------------------------------------------------------------------
// Evaluate one unit: residual rows, error contribution, and the evaluation record consumed by its
// Jacobian cells. kJacobian=false is the getError() path (joint_error_function-inl.h:35-54 etc.):
// same value, no residual/record stores.
// Returns the unit's error contribution (float, summed in double by the caller like the reference).
// ------------------------------------------------------------------------------------------------
template <bool kJacobian>
MB2_HD float evalUnit(const FunctionTables& T, int ui, const float* theta, const float* jp, const float* js, const float* targets,
                      const float* cweights, float* rec, float* residual) {
  const UnitDesc& u = T.units[ui];
  const EfDesc& e = T.efs[u.ef];
  float* r = kJacobian ? residual + u.row0 : nullptr;
  float* rc = kJacobian ? rec + u.recOff : nullptr;
  if (u.pad[0] != 0) { // limit gated off by enabledParameters_/activeJointParams_: zero row(s), no error
    if (kJacobian) { for (int k = 0; k < u.numRows; ++k) r[k] = 0.f; rc[0] = 0.f; }
    return 0.f;
  }
  switch (u.kind) {
    case kUnitPosition: {
      const float cw = cweights[u.weightIdx];
      if (cw == 0.f) { // joint_error_function-inl.h:197-199
        if (kJacobian) { r[0] = r[1] = r[2] = 0.f; rc[0] = rc[1] = rc[2] = rc[3] = 0.f; }
        return 0.f;
      }
      const float* ps = js + u.joint * kJointStateStride;
      const float* tg = targets + u.targetOff;
      // the constraint offset is shared by the batch (u.f) or, for an "instanced" block, follows the target in the instance's record
      const F3 off = u.pad[2] != 0 ? f3(tg[3], tg[4], tg[5]) : f3(u.f[0], u.f[1], u.f[2]);
      const F3 v = ld3(ps) + qrot(ld4(ps + 3), ps[7] * off); // transform.h:193-195
      const F3 f = f3(v.x - tg[0], v.y - tg[1], v.z - tg[2]);
      const float sq = dot(f, f);
      if (!kJacobian) return cw * lossValue(e, sq) * e.weight;
      const float w = cw * e.weight;
      float ds = sqrtf(w * lossDeriv(e, sq));
      r[0] = ds * f.x; r[1] = ds * f.y; r[2] = ds * f.z;
      if (fabsf(ds) < 1e-9f) ds = 0.f; // :216-218 rows stay zero
      rc[0] = v.x; rc[1] = v.y; rc[2] = v.z; rc[3] = ds;
      return w * lossValue(e, sq);
    }
    case kUnitOrientation:
    case kUnitOrientationRotDiff: {
      const float cw = cweights[u.weightIdx];
      if (cw == 0.f) {
        if (kJacobian) { for (int k = 0; k < 9; ++k) r[k] = 0.f; for (int k = 0; k < 10; ++k) rc[k] = 0.f; }
        return 0.f;
      }
      const Q4 q = ld4(js + u.joint * kJointStateStride + 3);
      float ro[9], rt[9], f[9], v[9];
      qmat(q4(u.f[0], u.f[1], u.f[2], u.f[3]), ro);
      qmat(ld4(targets + u.targetOff), rt);
      for (int k = 0; k < 3; ++k) {
        const F3 vk = qrot(q, f3(ro[3 * k], ro[3 * k + 1], ro[3 * k + 2]));
        v[3 * k] = vk.x; v[3 * k + 1] = vk.y; v[3 * k + 2] = vk.z;
      }
      if (u.kind == kUnitOrientation) {
        for (int k = 0; k < 9; ++k) f[k] = v[k] - rt[k];
      } else {
        // NB: the reference forms vec = R(q) * R(offset) as a matrix product; v_k = vec.col(k) (orientation_error_function.cpp:52-60)
        float rq[9];
        qmat(q, rq);
        for (int k = 0; k < 3; ++k)
          for (int rr = 0; rr < 3; ++rr) v[3 * k + rr] = rq[rr] * ro[3 * k] + rq[3 + rr] * ro[3 * k + 1] + rq[6 + rr] * ro[3 * k + 2];
        for (int k = 0; k < 3; ++k)
          for (int rr = 0; rr < 3; ++rr) // (R_t^T vec)(rr,k) - I
            f[3 * k + rr] = rt[3 * rr] * v[3 * k] + rt[3 * rr + 1] * v[3 * k + 1] + rt[3 * rr + 2] * v[3 * k + 2] - (rr == k ? 1.f : 0.f);
      }
      float sq = 0.f;
      for (int k = 0; k < 9; ++k) sq += f[k] * f[k];
      if (!kJacobian) return cw * lossValue(e, sq) * e.weight;
      const float w = cw * e.weight;
      float ds = sqrtf(w * lossDeriv(e, sq));
      for (int k = 0; k < 9; ++k) r[k] = ds * f[k];
      if (fabsf(ds) < 1e-9f) ds = 0.f;
      for (int k = 0; k < 9; ++k) rc[k] = v[k];
      rc[9] = ds;
      return w * lossValue(e, sq);
    }
    case kUnitStateMatrix:
    case kUnitStateLogMap: {
      const float* ps = js + u.joint * kJointStateStride;
      const float* tg = targets + u.targetOff; // t(3) q(4) s
      const F3 td = f3(ps[0] - tg[0], ps[1] - tg[1], ps[2] - tg[2]);
      const Q4 target = qnormalized(ld4(tg + 3));
      const Q4 rot = ld4(ps + 3);
      const float posW = u.f[0], rotW = u.f[1];
      float rotationError = 0.f;
      F3 lv = f3(0, 0, 0);
      float rd[9];
      if (u.kind == kUnitStateLogMap) {
        lv = quaternionLogMap(qmul(qconj(target), rot));
        rotationError = dot(lv, lv);
      } else {
        float a[9], b[9];
        qmat(rot, a); qmat(target, b);
        for (int k = 0; k < 9; ++k) { rd[k] = a[k] - b[k]; rotationError += rd[k] * rd[k]; }
      }
      if (!kJacobian) { // state_error_function.cpp:251-255
        float err = rotationError * kStateOrientationWeight * e.rotWgt * rotW;
        err += dot(td, td) * kStatePositionWeight * e.posWgt * posW;
        return err * e.weight;
      }
      const float pwgt = kStatePositionWeight * e.posWgt * e.weight * posW; // :440-441
      const float rwgt = kStateOrientationWeight * e.rotWgt * e.weight * rotW;
      const float wgt = sqrtf(pwgt), awgt = sqrtf(rwgt);
      r[0] = td.x * wgt; r[1] = td.y * wgt; r[2] = td.z * wgt;
      rc[0] = wgt; rc[1] = awgt;
      if (u.kind == kUnitStateLogMap) {
        r[3] = lv.x * awgt; r[4] = lv.y * awgt; r[5] = lv.z * awgt;
        quaternionLogMapDerivative(qmul(qconj(target), rot), rc + 2);
      } else {
        for (int k = 0; k < 9; ++k) r[3 + k] = rd[k] * awgt;
      }
      return dot(td, td) * pwgt + rotationError * rwgt;
    }
    case kUnitPlane: { // plane_error_function.cpp:49-70 through joint_error_function-inl.h (FuncDim = 1)
      const float cw = cweights[u.weightIdx];
      if (cw == 0.f) {
        if (kJacobian) { r[0] = 0.f; for (int k = 0; k < 6; ++k) rc[k] = 0.f; }
        return 0.f;
      }
      const float* ps = js + u.joint * kJointStateStride;
      const F3 v = ld3(ps) + qrot(ld4(ps + 3), ps[7] * f3(u.f[0], u.f[1], u.f[2]));
      const float* tg = targets + u.targetOff; // normal (not necessarily unit: PlaneDataT's ctor normalises), d
      const float nlen = sqrtf(tg[0] * tg[0] + tg[1] * tg[1] + tg[2] * tg[2]);
      const F3 nrm = f3(tg[0] / nlen, tg[1] / nlen, tg[2] / nlen);
      float val = dot(v, nrm) - tg[3];
      if (e.halfPlane && val > 0.f) val = 0.f;
      const bool on = !e.halfPlane || val < 0.f;
      const float sq = val * val;
      if (!kJacobian) return cw * lossValue(e, sq) * e.weight;
      const float w = cw * e.weight;
      float ds = sqrtf(w * lossDeriv(e, sq));
      r[0] = ds * val;
      if (fabsf(ds) < 1e-9f || !on) ds = 0.f; // :216-223 tiny scale or all-zero dfdv: the row stays zero
      rc[0] = v.x; rc[1] = v.y; rc[2] = v.z;
      rc[3] = ds * nrm.x; rc[4] = ds * nrm.y; rc[5] = ds * nrm.z;
      return w * lossValue(e, sq);
    }
    case kUnitModelParameter: { // model_parameters_error_function.cpp:38-58, 90-133; kMotionWeight = 1e-1 (.h:61)
      const float pdiff = u.f[0] * (theta[u.i[0]] - targets[u.targetOff]);
      const float scale = e.weight * 1e-1f;
      if (u.pad[1] != 0) return kJacobian ? 0.f : pdiff * pdiff * scale; // negative target weight: counted by getError only (:56-59 vs :113)
      if (kJacobian) { const float sw = sqrtf(scale); r[0] = pdiff * sw; rc[0] = sw; }
      return pdiff * pdiff * scale;
    }
    default: break;
  }
  // ---- limits (limit_error_function.cpp) ----
  const bool L2 = e.lossType == kLossL2;
  const float lw = u.f[7]; // limit.weight
  if (!kJacobian) {
    // getErrorImpl (:818-867): per-limit term, then * kLimitWeight * weight (* invC2 for L2)
    const float post = kLimitWeight * e.weight * (L2 ? e.invC2 : 1.f);
    float sq = -1.f;
    float pre = lw;
    switch (u.kind) {
      case kUnitLimitMinMax: {
        const float p = theta[u.i[0]];
        if (p < u.f[0]) sq = (u.f[0] - p) * (u.f[0] - p);
        if (p > u.f[1]) sq = (u.f[1] - p) * (u.f[1] - p);
        break;
      }
      case kUnitLimitMinMaxJoint: {
        const float p = jointParameterAt(T, jp, theta, u.i[0]);
        if (p < u.f[0]) sq = (u.f[0] - p) * (u.f[0] - p);
        if (p > u.f[1]) sq = (u.f[1] - p) * (u.f[1] - p);
        break;
      }
      case kUnitLimitLinear: {
        if (limitInRange(u.f[2], u.f[3], theta[u.i[1]])) { const float res = theta[u.i[1]] * u.f[0] - u.f[1] - theta[u.i[0]]; sq = res * res; }
        break;
      }
      case kUnitLimitLinearJoint: {
        const float p1 = jointParameterAt(T, jp, theta, u.i[1]);
        if (limitInRange(u.f[2], u.f[3], p1)) { const float res = p1 * u.f[0] - u.f[1] - jointParameterAt(T, jp, theta, u.i[0]); sq = res * res; }
        break;
      }
      case kUnitLimitHalfPlane: {
        const float res = theta[u.i[0]] * u.f[0] + theta[u.i[1]] * u.f[1] - u.f[2];
        if (res < 0.f) sq = res * res;
        break;
      }
      case kUnitLimitEllipsoid: {
        F3 pos, diff;
        evalEllipsoid(T, u, js, pos, diff);
        sq = dot(diff, diff);
        pre = kLimitPositionWeight * lw;
        break;
      }
      default: break;
    }
    if (sq < 0.f) return 0.f;
    return pre * (L2 ? sq : lossValue(e, sq)) * post;
  }
  // Jacobian path (:992-1121): tWeight folds invC2 for L2
  const float tWeight = kLimitWeight * e.weight * (L2 ? e.invC2 : 1.f);
  float res = 0.f;
  bool active = false;
  switch (u.kind) {
    case kUnitLimitMinMax: {
      const float p = theta[u.i[0]];
      if (p < u.f[0]) { res = p - u.f[0]; active = true; }
      else if (p > u.f[1]) { res = p - u.f[1]; active = true; }
      break;
    }
    case kUnitLimitMinMaxJoint: {
      const float p = jointParameterAt(T, jp, theta, u.i[0]);
      if (p < u.f[0]) { res = p - u.f[0]; active = true; }
      else if (p > u.f[1]) { res = p - u.f[1]; active = true; }
      break;
    }
    case kUnitLimitLinear:
      if (limitInRange(u.f[2], u.f[3], theta[u.i[1]])) { res = theta[u.i[1]] * u.f[0] - u.f[1] - theta[u.i[0]]; active = true; }
      break;
    case kUnitLimitLinearJoint:
      { const float p1 = jointParameterAt(T, jp, theta, u.i[1]);
        if (limitInRange(u.f[2], u.f[3], p1)) { res = p1 * u.f[0] - u.f[1] - jointParameterAt(T, jp, theta, u.i[0]); active = true; } }
      break;
    case kUnitLimitHalfPlane:
      res = theta[u.i[0]] * u.f[0] + theta[u.i[1]] * u.f[1] - u.f[2];
      active = res < 0.f;
      break;
    case kUnitLimitEllipsoid: {
      F3 pos, diff;
      evalEllipsoid(T, u, js, pos, diff);
      const float sq = dot(diff, diff);
      const float jwgt = L2 ? sqrtf(tWeight * kLimitPositionWeight * lw) : sqrtf(tWeight * kLimitPositionWeight * lw * lossDeriv(e, sq));
      r[0] = diff.x * jwgt; r[1] = diff.y * jwgt; r[2] = diff.z * jwgt;
      rc[0] = pos.x; rc[1] = pos.y; rc[2] = pos.z; rc[3] = jwgt;
      return L2 ? tWeight * kLimitPositionWeight * lw * sq : tWeight * kLimitPositionWeight * lw * lossValue(e, sq);
    }
    default: break;
  }
  if (!active) { r[0] = 0.f; rc[0] = 0.f; return 0.f; }
  const float sq = res * res;
  const float wl = L2 ? sqrtf(tWeight * lw) : sqrtf(tWeight * lw * lossDeriv(e, sq));
  r[0] = res * wl;
  rc[0] = wl;
  return L2 ? tWeight * lw * sq : tWeight * lw * lossValue(e, sq);
}

// ------------------------------------------------------------------------------------------------
// Fill one Jacobian cell: rows [row0, row0+numRows) of column cell.col. Jcol points at the start
// of that column (row stride 1). Contributions are summed in the reference's walk order.
// ------------------------------------------------------------------------------------------------
MB2_HD void jacobianCell(const FunctionTables& T, int ci, const float* js, const float* rec, const float* targets, float* Jbase) {
  const CellDesc& c = T.cells[ci];
  const UnitDesc& u = T.units[c.unit];
  const float* rc = rec + u.recOff;
  // row k of the unit lands at out[MB2_ROW(k)]: contiguous in the K-major matrix; in the strip layout four rows of one quad are
  // contiguous and consecutive quads are quadStride strips apart
  float* out = T.stripMode ? Jbase + c.stripOff : Jbase + (size_t)c.col * T.ldJ + u.row0;
  const int rowShift = T.stripMode ? (u.row0 & 3) : 0, quadFloats = T.stripMode ? int(c.quadStride) * 64 : 4;
#define MB2_ROW(k) ((((k) + rowShift) >> 2) * quadFloats + (((k) + rowShift) & 3))
  const ContribDesc* cb = T.contribs + c.contribBegin;
  switch (u.kind) {
    case kUnitPosition:
    case kUnitLimitEllipsoid: {
      const F3 v = ld3(rc);
      const float ds = rc[3];
      F3 acc = f3(0.f, 0.f, 0.f);
      for (int k = 0; k < c.contribCount; ++k) acc = acc + (pointDerivative(T, js, cb[k].joint, cb[k].dof, v) * ds) * cb[k].coef;
#if defined(__CUDA_ARCH__)
      // strip layout, unit on a quad boundary (every multi-row unit owns whole quads: its fourth row is a structural zero): the three
      // rows of this column are ONE 16-byte store instead of three scattered 4-byte ones (the global stores were a third of the sweep's
      // LSU wavefronts, its busiest pipe)
      if (T.stripMode && rowShift == 0) { *reinterpret_cast<float4*>(out) = make_float4(acc.x, acc.y, acc.z, 0.f); break; }
#endif
      out[MB2_ROW(0)] = acc.x; out[MB2_ROW(1)] = acc.y; out[MB2_ROW(2)] = acc.z;
      break;
    }
    case kUnitPlane: {
      const F3 v = ld3(rc);
      F3 acc = f3(0.f, 0.f, 0.f);
      for (int k = 0; k < c.contribCount; ++k) acc = acc + pointDerivative(T, js, cb[k].joint, cb[k].dof, v) * cb[k].coef;
      out[MB2_ROW(0)] = dot(ld3(rc + 3), acc); // (sqrt(w loss') n)^T dv/dp
      break;
    }
    case kUnitOrientation:
    case kUnitOrientationRotDiff: {
      const float ds = rc[9];
      float acc[9];
      for (int k = 0; k < 9; ++k) acc[k] = 0.f;
      float it[9]; // R_target (RotDiff: rows 3k.. = R_t^T * d)
      if (u.kind == kUnitOrientationRotDiff) qmat(ld4(targets + u.targetOff), it);
      for (int k = 0; k < c.contribCount; ++k) {
        const F3 axis = rotationAxisCol(js, cb[k].joint, cb[k].dof - 3);
        for (int jv = 0; jv < 3; ++jv) {
          F3 d = cross(axis, ld3(rc + 3 * jv));
          if (u.kind == kUnitOrientationRotDiff) d = f3(it[0] * d.x + it[1] * d.y + it[2] * d.z, it[3] * d.x + it[4] * d.y + it[5] * d.z, it[6] * d.x + it[7] * d.y + it[8] * d.z);
          acc[3 * jv] += (ds * d.x) * cb[k].coef; acc[3 * jv + 1] += (ds * d.y) * cb[k].coef; acc[3 * jv + 2] += (ds * d.z) * cb[k].coef;
        }
      }
#if defined(__CUDA_ARCH__)
      if (T.stripMode && rowShift == 0) { // nine rows = three quads of this column (the last one holds row 8 and three structural zeros)
        *reinterpret_cast<float4*>(out) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(out + quadFloats) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        *reinterpret_cast<float4*>(out + 2 * quadFloats) = make_float4(acc[8], 0.f, 0.f, 0.f);
        break;
      }
#endif
      for (int k = 0; k < 9; ++k) out[MB2_ROW(k)] = acc[k];
      break;
    }
    case kUnitStateMatrix:
    case kUnitStateLogMap: {
      const float wgt = rc[0], awgt = rc[1];
      const float* ps = js + u.joint * kJointStateStride;
      const F3 ti = ld3(ps);
      const Q4 rot = ld4(ps + 3);
      const bool lm = u.kind == kUnitStateLogMap;
      float acc[12];
      for (int k = 0; k < 12; ++k) acc[k] = 0.f;
      float rm[9];
      Q4 target = q4(0, 0, 0, 1);
      if (lm) target = qnormalized(ld4(targets + u.targetOff + 3)); else qmat(rot, rm);
      for (int k = 0; k < c.contribCount; ++k) {
        const int a = cb[k].joint, d = cb[k].dof;
        const float coef = cb[k].coef;
        const F3 jc = pointDerivative(T, js, a, d, ti) * wgt; // state_error_function.cpp:498-510,546-550
        acc[0] += jc.x * coef; acc[1] += jc.y * coef; acc[2] += jc.z * coef;
        if (d >= 3 && d < 6) {
          const F3 axis = rotationAxisCol(js, a, d - 3);
          if (lm) {
            const F3 jr = logMapRelativeDerivativeQ1(rot, target, axis, rc + 2) * awgt;
            acc[3] += jr.x * coef; acc[4] += jr.y * coef; acc[5] += jr.z * coef;
          } else { // vec([axis]x R) * awgt (state_error_function.cpp:116-121)
            for (int cc = 0; cc < 3; ++cc) {
              const F3 col = cross(axis, f3(rm[3 * cc], rm[3 * cc + 1], rm[3 * cc + 2]));
              acc[3 + 3 * cc] += (col.x * awgt) * coef; acc[4 + 3 * cc] += (col.y * awgt) * coef; acc[5 + 3 * cc] += (col.z * awgt) * coef;
            }
          }
        }
      }
      const int nr = lm ? 6 : 12;
      for (int k = 0; k < nr; ++k) out[MB2_ROW(k)] = acc[k];
      break;
    }
    default: // simple limits: value = wgtLoss * static coefficient
      out[MB2_ROW(0)] = rc[0] * c.coef;
      break;
  }
#undef MB2_ROW
}

} // namespace mb2

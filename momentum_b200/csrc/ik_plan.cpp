#include "ik_plan.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <map>

namespace mb2 {

std::string HostCharacter::validate() const {
  const int J = numJoints;
  if (J <= 0) return "skeleton has no joints";
  if (numParams <= 0 || numParams > 2048) return "number of model parameters must be in [1, 2048] (ParameterSet is bitset<2048>)";
  if (int(parent.size()) != J || int(offset.size()) != 3 * J || int(prerot.size()) != 4 * J) return "joint array sizes do not match the joint count";
  for (int j = 0; j < J; ++j)
    if (parent[j] >= j || parent[j] < -1) return "skeleton is not topologically sorted: parent index must precede child (skeleton.h:23-24)";
  const int rows = J * kParametersPerJoint;
  if (int(ptOuter.size()) != rows + 1 || int(ptOffsets.size()) != rows) return "parameter transform must have 7 * numJoints rows";
  if (ptOuter[0] != 0) return "parameter transform outer index must start at 0";
  for (int r = 0; r < rows; ++r)
    if (ptOuter[r + 1] < ptOuter[r]) return "parameter transform outer index must be non-decreasing";
  if (ptOuter[rows] != int(ptInner.size()) || ptInner.size() != ptVals.size()) return "parameter transform nnz mismatch";
  for (int c : ptInner)
    if (c < 0 || c >= numParams) return "parameter transform column out of range";
  return "";
}

void HostCharacter::buildLevels() {
  std::vector<int> depth(numJoints, 0);
  int maxDepth = 0;
  for (int j = 0; j < numJoints; ++j) {
    depth[j] = parent[j] < 0 ? 0 : depth[parent[j]] + 1;
    maxDepth = std::max(maxDepth, depth[j]);
  }
  levelStart.assign(maxDepth + 2, 0);
  for (int j = 0; j < numJoints; ++j) levelStart[depth[j] + 1]++;
  for (int l = 0; l <= maxDepth; ++l) levelStart[l + 1] += levelStart[l];
  levelJoints.resize(numJoints);
  std::vector<int> cursor(levelStart.begin(), levelStart.end() - 1);
  for (int j = 0; j < numJoints; ++j) levelJoints[cursor[depth[j]]++] = j;
}

std::vector<uint8_t> HostCharacter::computeActiveJointParams(const std::vector<uint8_t>& enabled) const {
  std::vector<uint8_t> r(size_t(numJoints) * kParametersPerJoint, 0);
  for (int row = 0; row < numJoints * kParametersPerJoint; ++row)
    for (int k = ptOuter[row]; k < ptOuter[row + 1]; ++k)
      if (enabled[ptInner[k]]) r[row] = 1;
  return r;
}

int32_t jacobianBlockSize(const HostCharacter& ch, const HostErrorFunction& ef) {
  switch (ef.kind) {
    case 0: return 3 * ef.numConstraints();
    case 1:
    case 2: return 9 * ef.numConstraints();
    case 3: {
      int n = 0;
      for (int j = 0; j < ch.numJoints; ++j) n += (ef.posW[j] != 0.f || ef.rotW[j] != 0.f) ? 1 : 0;
      return n * (ef.rotationErrorType == 1 ? 6 : 12);
    }
    case 5: return ef.numConstraints(); // joint_error_function-inl.h:300-302 with FuncDim = 1
    case 6: { // model_parameters_error_function.cpp:93-95
      int n = 0;
      for (float w : ef.paramWeights) n += w > 0.f ? 1 : 0;
      return n;
    }
    case 4: {
      int n = 0;
      for (const auto& l : ch.limits) {
        if (l.type == 2) continue; // MinMaxJointPassive
        n += (l.type == 5) ? 3 : 1;
      }
      return n;
    }
  }
  return 0;
}

static EfDesc makeEfDesc(const HostErrorFunction& ef) {
  EfDesc d{};
  d.weight = ef.weight;
  d.alpha = ef.lossAlpha;
  d.invC2 = 1.f / (ef.lossC * ef.lossC);
  // GeneralizedLossT ctor snapping (math/generalized_loss.cpp:81-101), kEps = 1e-9
  const float kEps = 1e-9f;
  const float a = ef.lossAlpha;
  if (a >= 2.f - kEps && a <= 2.f + kEps) d.lossType = kLossL2;
  else if (a >= 1.f - kEps && a <= 1.f + kEps) d.lossType = kLossL1;
  else if (a >= 0.f - kEps && a <= 0.f + kEps) d.lossType = kLossCauchy;
  else if (a == -FLT_MAX || (std::isinf(a) && a < 0)) d.lossType = kLossWelsch;
  else d.lossType = kLossGeneral;
  d.posWgt = ef.posWgt;
  d.rotWgt = ef.rotWgt;
  d.kind = ef.kind;
  d.halfPlane = ef.halfPlane ? 1 : 0;
  return d;
}

namespace {
struct CellBuilder {
  // column -> contributions in walk order
  std::map<int, std::vector<ContribDesc>> m;
  void add(const HostCharacter& ch, int jointParam, int joint, int dof, const std::vector<uint8_t>* enabledGate) {
    for (int k = ch.ptOuter[jointParam]; k < ch.ptOuter[jointParam + 1]; ++k) {
      const int col = ch.ptInner[k];
      if (enabledGate != nullptr && !(*enabledGate)[col]) continue;
      ContribDesc c;
      c.joint = uint16_t(joint);
      c.dof = uint16_t(dof);
      c.coef = ch.ptVals[k];
      m[col].push_back(c);
    }
  }
};
} // namespace

std::string buildPlan(const HostCharacter& ch, const std::vector<HostErrorFunction>& efs, const std::vector<uint8_t>& enabled, bool compact, Plan& out,
                      const std::vector<int32_t>* columnOrder, bool alignRowGroups) {
  out = Plan();
  out.compact = compact;
  const int n = ch.numParams;
  if (int(enabled.size()) != n) return "enabled parameter set size mismatch";
  for (int i = 0; i < n; ++i)
    if (enabled[i]) { out.actualParameters = i + 1; out.enabledList.push_back(i); }
  const std::vector<uint8_t> active = ch.computeActiveJointParams(enabled);
  std::vector<int> colMap(n, -1); // model parameter -> device column
  if (compact) {
    out.deviceCols = out.enabledList;
    if (columnOrder != nullptr) out.deviceCols = *columnOrder; // may hold -1 entries: all-zero alignment columns (ik_chol_sched.h)
    size_t listed = 0;
    for (size_t a = 0; a < out.deviceCols.size(); ++a) {
      const int p = out.deviceCols[a];
      if (p == -1 && columnOrder != nullptr) continue;
      if (p < 0 || p >= n || !enabled[p] || colMap[p] >= 0) return "column order must list every enabled parameter once";
      colMap[p] = int(a);
      ++listed;
    }
    if (listed != out.enabledList.size()) return "column order must list every enabled parameter once";
    out.numCols = int(out.deviceCols.size());
  } else {
    for (int i = 0; i < n; ++i) colMap[i] = i;
    out.numCols = n;
    out.deviceCols.resize(n);
    for (int i = 0; i < n; ++i) out.deviceCols[i] = i;
  }

  int row = 0, rec = 0;
  auto flushCells = [&](int unitIndex, CellBuilder& cb) {
    for (auto& kv : cb.m) {
      if (colMap[kv.first] < 0) continue; // column of a disabled parameter: not held on the device
      CellDesc c{};
      c.unit = uint16_t(unitIndex);
      c.col = uint16_t(colMap[kv.first]);
      c.contribBegin = uint32_t(out.contribs.size());
      c.contribCount = uint16_t(kv.second.size());
      c.coef = 0.f;
      out.contribs.insert(out.contribs.end(), kv.second.begin(), kv.second.end());
      out.cells.push_back(c);
    }
  };
  auto staticCell = [&](int unitIndex, int col, float coef) {
    if (colMap[col] < 0) return;
    CellDesc c{};
    c.unit = uint16_t(unitIndex);
    c.col = uint16_t(colMap[col]);
    c.contribBegin = 0;
    c.contribCount = 0;
    c.coef = coef;
    out.cells.push_back(c);
  };

  for (size_t e = 0; e < efs.size(); ++e) {
    const HostErrorFunction& ef = efs[e];
    out.efs.push_back(makeEfDesc(ef));
    if (!(ef.weight > 0.f)) continue; // skeleton_solver_function.cpp:228-230: disabled block has no rows
    if (ef.kind <= 2) {
      const bool isPos = ef.kind == 0;
      const int per = isPos ? 3 : 4;
      for (int c = 0; c < ef.numConstraints(); ++c) {
        if (ef.parents[c] < 0 || ef.parents[c] >= ch.numJoints) return "constraint parent joint out of range";
        UnitDesc u{};
        u.kind = isPos ? kUnitPosition : (ef.kind == 1 ? kUnitOrientation : kUnitOrientationRotDiff);
        u.ef = int32_t(e);
        u.joint = ef.parents[c];
        if (alignRowGroups) row = (row + 3) & ~3;
        u.row0 = row;
        u.numRows = isPos ? 3 : 9;
        const bool instanced = isPos && ef.instanceOffsets;
        u.targetOff = ef.targetOff + (instanced ? 6 : per) * c;
        u.weightIdx = ef.weightOff + c;
        u.recOff = rec;
        u.extra = -1;
        u.pad[2] = instanced ? 1 : 0;
        if (!instanced) for (int k = 0; k < per; ++k) u.f[k] = ef.offsets[size_t(per) * c + k];
        const int ui = int(out.units.size());
        out.units.push_back(u);
        row += u.numRows;
        if (alignRowGroups) row = (row + 3) & ~3; // a multi-row unit owns its row quads (one-row units pack among themselves)
        rec += isPos ? 4 : 10;
        CellBuilder cb;
        for (int jnt = ef.parents[c]; jnt >= 0; jnt = ch.parent[jnt]) { // joint_error_function-inl.h:229-294
          const int pb = jnt * kParametersPerJoint;
          if (isPos)
            for (int d = 0; d < 3; ++d)
              if (active[pb + d]) cb.add(ch, pb + d, jnt, d, &enabled);
          for (int d = 0; d < 3; ++d)
            if (active[pb + 3 + d]) cb.add(ch, pb + 3 + d, jnt, 3 + d, &enabled);
          if (isPos && active[pb + 6]) cb.add(ch, pb + 6, jnt, 6, &enabled);
        }
        flushCells(ui, cb);
      }
    } else if (ef.kind == 3) {
      const bool lm = ef.rotationErrorType == 1;
      for (int i = 0; i < ch.numJoints; ++i) {
        if (ef.rotW[i] == 0.f && ef.posW[i] == 0.f) continue; // state_error_function.cpp:424-426
        UnitDesc u{};
        u.kind = lm ? kUnitStateLogMap : kUnitStateMatrix;
        u.ef = int32_t(e);
        u.joint = i;
        if (alignRowGroups) row = (row + 3) & ~3;
        u.row0 = row;
        u.numRows = lm ? 6 : 12;
        u.targetOff = ef.targetOff + 8 * i;
        u.weightIdx = -1;
        u.recOff = rec;
        u.extra = -1;
        u.f[0] = ef.posW[i];
        u.f[1] = ef.rotW[i];
        const int ui = int(out.units.size());
        out.units.push_back(u);
        row += u.numRows;
        if (alignRowGroups) row = (row + 3) & ~3;
        rec += lm ? 14 : 2;
        CellBuilder cb;
        for (int jnt = i; jnt >= 0; jnt = ch.parent[jnt]) { // state_error_function.cpp:486-555 (no enabledParameters gate)
          const int pb = jnt * kParametersPerJoint;
          for (int d = 0; d < 3; ++d) {
            if (active[pb + d]) cb.add(ch, pb + d, jnt, d, nullptr);
            if (active[pb + 3 + d]) cb.add(ch, pb + 3 + d, jnt, 3 + d, nullptr);
          }
          if (active[pb + 6]) cb.add(ch, pb + 6, jnt, 6, nullptr);
        }
        flushCells(ui, cb);
      }
    } else if (ef.kind == 5) { // Plane: a one-row position-type constraint (JointErrorFunctionT<T, PlaneDataT<T>, 1>)
      for (int c = 0; c < ef.numConstraints(); ++c) {
        if (ef.parents[c] < 0 || ef.parents[c] >= ch.numJoints) return "constraint parent joint out of range";
        UnitDesc u{};
        u.kind = kUnitPlane;
        u.ef = int32_t(e);
        u.joint = ef.parents[c];
        u.row0 = row;
        u.numRows = 1;
        u.targetOff = ef.targetOff + 4 * c;
        u.weightIdx = ef.weightOff + c;
        u.recOff = rec;
        u.extra = -1;
        for (int k = 0; k < 3; ++k) u.f[k] = ef.offsets[size_t(3) * c + k];
        const int ui = int(out.units.size());
        out.units.push_back(u);
        row += 1;
        rec += 6; // world point, scaled normal
        CellBuilder cb;
        for (int jnt = ef.parents[c]; jnt >= 0; jnt = ch.parent[jnt]) { // joint_error_function-inl.h:229-294, NumPos = 1
          const int pb = jnt * kParametersPerJoint;
          for (int d = 0; d < 3; ++d)
            if (active[pb + d]) cb.add(ch, pb + d, jnt, d, &enabled);
          for (int d = 0; d < 3; ++d)
            if (active[pb + 3 + d]) cb.add(ch, pb + 3 + d, jnt, 3 + d, &enabled);
          if (active[pb + 6]) cb.add(ch, pb + 6, jnt, 6, &enabled);
        }
        flushCells(ui, cb);
      }
    } else if (ef.kind == 6) { // ModelParameters: rows are packed over the enabled parameters with weight > 0 (:113-123); the block
      if (int(ef.paramWeights.size()) != n) return "model-parameter target weights must have one entry per parameter"; // keeps getJacobianSize rows
      const int blockStart = row;
      for (int i = 0; i < n; ++i) {
        if (!enabled[i] || ef.paramWeights[i] == 0.f) continue;
        if (ef.paramWeights[i] < 0.f) { // no Jacobian row (:113 tests weight > 0) but getError still counts it (:56-59): an error-only unit
          UnitDesc u{};
          u.kind = kUnitModelParameter;
          u.ef = int32_t(e);
          u.joint = -1;
          u.row0 = row;
          u.numRows = 0;
          u.targetOff = ef.targetOff + i;
          u.weightIdx = -1;
          u.recOff = rec;
          u.extra = -1;
          u.i[0] = i;
          u.f[0] = ef.paramWeights[i];
          u.pad[1] = 1;
          out.units.push_back(u);
          continue;
        }
        UnitDesc u{};
        u.kind = kUnitModelParameter;
        u.ef = int32_t(e);
        u.joint = -1;
        u.row0 = row;
        u.numRows = 1;
        u.targetOff = ef.targetOff + i;
        u.weightIdx = -1;
        u.recOff = rec;
        u.extra = -1;
        u.i[0] = i;
        u.f[0] = ef.paramWeights[i];
        const int ui = int(out.units.size());
        out.units.push_back(u);
        row += 1;
        rec += 1;
        staticCell(ui, i, ef.paramWeights[i]);
      }
      row = blockStart + jacobianBlockSize(ch, ef);
    } else if (ef.kind == 4) {
      for (const HostLimit& l : ch.limits) {
        if (l.type == 2) continue; // MinMaxJointPassive: no rows (limit_error_function.cpp:1051-1052)
        UnitDesc u{};
        u.ef = int32_t(e);
        u.row0 = row;
        u.numRows = 1;
        u.targetOff = -1;
        u.weightIdx = -1;
        u.recOff = rec;
        u.extra = -1;
        u.f[7] = l.weight;
        const int ui = int(out.units.size());
        bool disabled = false;
        switch (l.type) {
          case 0: { // MinMax :459-503
            u.kind = kUnitLimitMinMax;
            u.i[0] = l.i[0];
            u.f[0] = l.f[0]; u.f[1] = l.f[1];
            if (l.i[0] < 0 || l.i[0] >= n) return "MinMax limit parameter index out of range";
            disabled = !enabled[l.i[0]];
            out.units.push_back(u);
            if (!disabled) staticCell(ui, l.i[0], 1.f);
            break;
          }
          case 1: { // MinMaxJoint :505-558
            u.kind = kUnitLimitMinMaxJoint;
            const int jpi = l.i[0] * kParametersPerJoint + l.i[1];
            if (l.i[0] < 0 || l.i[0] >= ch.numJoints || l.i[1] < 0 || l.i[1] > 6) return "MinMaxJoint limit index out of range";
            u.i[0] = jpi;
            u.f[0] = l.f[0]; u.f[1] = l.f[1];
            disabled = !active[jpi];
            out.units.push_back(u);
            if (!disabled)
              for (int k = ch.ptOuter[jpi]; k < ch.ptOuter[jpi + 1]; ++k) staticCell(ui, ch.ptInner[k], ch.ptVals[k]);
            break;
          }
          case 3: { // Linear :560-598
            u.kind = kUnitLimitLinear;
            const int ref = l.i[0], tgt = l.i[1];
            if (ref < 0 || ref >= n || tgt < 0 || tgt >= n) return "Linear limit parameter index out of range";
            u.i[0] = ref; u.i[1] = tgt;
            for (int k = 0; k < 4; ++k) u.f[k] = l.f[k];
            disabled = !enabled[tgt] && !enabled[ref];
            out.units.push_back(u);
            if (!disabled) {
              // assignments, target first then reference: the later one wins if both name the same column
              if (enabled[tgt] && !(enabled[ref] && ref == tgt)) staticCell(ui, tgt, l.f[0]);
              if (enabled[ref]) staticCell(ui, ref, -1.f);
            }
            break;
          }
          case 4: { // LinearJoint :600-656
            u.kind = kUnitLimitLinearJoint;
            const int ri = l.i[0] * kParametersPerJoint + l.i[1];
            const int ti = l.i[2] * kParametersPerJoint + l.i[3];
            if (l.i[0] < 0 || l.i[0] >= ch.numJoints || l.i[2] < 0 || l.i[2] >= ch.numJoints || l.i[1] < 0 || l.i[1] > 6 || l.i[3] < 0 || l.i[3] > 6)
              return "LinearJoint limit index out of range";
            u.i[0] = ri; u.i[1] = ti;
            for (int k = 0; k < 4; ++k) u.f[k] = l.f[k];
            disabled = !active[ri] && !active[ti];
            out.units.push_back(u);
            if (!disabled) {
              std::map<int, float> acc;
              if (active[ti]) for (int k = ch.ptOuter[ti]; k < ch.ptOuter[ti + 1]; ++k) acc[ch.ptInner[k]] += l.f[0] * ch.ptVals[k];
              if (active[ri]) for (int k = ch.ptOuter[ri]; k < ch.ptOuter[ri + 1]; ++k) acc[ch.ptInner[k]] += -ch.ptVals[k];
              for (auto& kv : acc) staticCell(ui, kv.first, kv.second);
            }
            break;
          }
          case 6: { // HalfPlane :658-699
            u.kind = kUnitLimitHalfPlane;
            const int p1 = l.i[0], p2 = l.i[1];
            if (p1 < 0 || p1 >= n || p2 < 0 || p2 >= n) return "HalfPlane limit parameter index out of range";
            u.i[0] = p1; u.i[1] = p2;
            u.f[0] = l.f[0]; u.f[1] = l.f[1]; u.f[2] = l.f[2];
            disabled = !enabled[p1] && !enabled[p2];
            out.units.push_back(u);
            if (!disabled) {
              if (enabled[p1] && !(enabled[p2] && p1 == p2)) staticCell(ui, p1, l.f[0]);
              if (enabled[p2]) staticCell(ui, p2, l.f[1]);
            }
            break;
          }
          case 5: { // Ellipsoid :701-785
            u.kind = kUnitLimitEllipsoid;
            u.numRows = 3;
            if (alignRowGroups) { row = (row + 3) & ~3; u.row0 = row; }
            u.i[0] = l.i[0]; // ellipsoidParent
            u.joint = l.i[1]; // parent
            if (l.i[0] < 0 || l.i[0] >= ch.numJoints || l.i[1] < 0 || l.i[1] >= ch.numJoints) return "Ellipsoid limit joint index out of range";
            u.extra = int32_t(out.limitData.size());
            out.limitData.insert(out.limitData.end(), l.f, l.f + 27);
            out.limitData.push_back(0.f);
            out.units.push_back(u);
            CellBuilder cb;
            for (int jnt = l.i[1]; jnt != l.i[0] && jnt >= 0; jnt = ch.parent[jnt]) { // :740-777
              const int pb = jnt * kParametersPerJoint;
              for (int d = 0; d < 3; ++d) {
                if (active[pb + d]) cb.add(ch, pb + d, jnt, d, nullptr);
                if (active[pb + 3 + d]) cb.add(ch, pb + 3 + d, jnt, 3 + d, nullptr);
              }
              if (active[pb + 6]) cb.add(ch, pb + 6, jnt, 6, nullptr);
            }
            flushCells(ui, cb);
            break;
          }
          default: return "Unknown parameter type for joint limit";
        }
        out.units[ui].pad[0] = disabled ? 1 : 0;
        row += out.units[ui].numRows;
        if (alignRowGroups && out.units[ui].numRows > 1) row = (row + 3) & ~3;
        rec += (l.type == 5) ? 4 : 1;
      }
    } else {
      return "unknown error function kind";
    }
  }
  if (out.units.size() > 65535) return "too many constraints in one solver function (limit 65535 units)";
  out.numRows = row;
  out.recStride = std::max(rec, 1);
  // Group cells so that neighbouring lanes run the same code path and store next to each other. K-major matrix: same column, consecutive
  // units = consecutive rows of that column. Strip layout (alignRowGroups): same unit, consecutive device columns = consecutive 16-byte
  // pieces of the unit's strip (a row quad x 16 columns is 256 contiguous bytes), so a warp's stores fill whole sectors and lines, and
  // the unit / record reads of a warp are broadcasts.
  std::stable_sort(out.cells.begin(), out.cells.end(), [&](const CellDesc& a, const CellDesc& b) {
    const int ka = out.units[a.unit].kind, kb = out.units[b.unit].kind;
    if (ka != kb) return ka < kb;
    if (alignRowGroups) return a.unit != b.unit ? a.unit < b.unit : a.col < b.col;
    if (a.col != b.col) return a.col < b.col;
    return a.unit < b.unit;
  });
  return "";
}

} // namespace mb2

// Host-side model of a character and of the batched solver function's objective, and the planner
// that flattens it into the unit / cell / contribution tables consumed by the device sweep.
//
// The planner performs, once per (constraint topology, enabled-parameter set), the ancestor walks
// that the reference repeats for every constraint on every iteration
// (joint_error_function-inl.h:229-294, state_error_function.cpp:486-555, limit_error_function.cpp:740-777),
// including its gating rules (activeJointParams_, enabledParameters_, zero weights).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "ik_types.h"

namespace mb2 {

struct HostLimit { // character/parameter_limits.h:117-127, flattened like mb2_parameter_limit
  int32_t type;
  float weight;
  int32_t i[4];
  float f[27];
};

struct HostCharacter {
  int32_t numJoints{0}, numParams{0};
  std::vector<int32_t> parent;
  std::vector<float> offset, prerot;
  std::vector<int32_t> ptOuter, ptInner;
  std::vector<float> ptVals, ptOffsets;
  std::vector<HostLimit> limits;
  // derived
  std::vector<int32_t> levelStart, levelJoints;
  std::string validate() const; // empty when fine (MT_CHECK-style message otherwise)
  void buildLevels();
  // ParameterTransformT::computeActiveJointParams (parameter_transform.cpp:97-107)
  std::vector<uint8_t> computeActiveJointParams(const std::vector<uint8_t>& enabled) const;
};

struct HostErrorFunction {
  int32_t kind{0}; // 0 position, 1 orientation, 2 orientation rot-diff, 3 state, 4 limit, 5 plane, 6 model parameters
  float weight{1.f};
  float lossAlpha{2.f}, lossC{1.f};
  std::vector<int32_t> parents;
  std::vector<float> offsets; // 3 or 4 per constraint
  std::vector<float> weights; // shared constraint weights
  int32_t rotationErrorType{0};
  float posWgt{1.f}, rotWgt{1.f};
  std::vector<float> posW, rotW;
  bool halfPlane{false};            // plane: PlaneErrorFunctionT(above)
  bool instanceOffsets{false};      // position: offsets are per instance (record = target xyz, offset xyz per constraint)
  std::vector<float> paramWeights; // model parameters: targetWeights_ [numParams]
  // layout (assigned when added)
  int32_t targetOff{0}, targetSize{0}; // floats per instance
  int32_t weightOff{0};                // into the constraint-weight array
  int32_t numConstraints() const { return int32_t(parents.size()); }
};

struct Plan {
  std::vector<EfDesc> efs;
  std::vector<UnitDesc> units;
  std::vector<CellDesc> cells;
  std::vector<ContribDesc> contribs;
  std::vector<float> limitData;
  int32_t numRows{0};   // m, unpadded (sum of getJacobianBlockSize of blocks with weight > 0)
  int32_t recStride{0};
  int32_t actualParameters{0}; // skeleton_solver_function.cpp:45-52
  std::vector<int32_t> enabledList; // gauss_newton_solver.cpp:57-66
  int32_t numCols{0};   // device Jacobian columns: numParams (full) or enabledList.size() (compact)
  bool compact{false};
  std::vector<int32_t> deviceCols; // [numCols] model parameter held by each device column
};

// Builds the plan. `enabled` has numParams entries.
// compact = true drops the columns of disabled parameters and packs the enabled ones in order (what the solver
// needs: gauss_newton_solver.cpp:204-209 discards the others anyway); compact = false keeps the reference's
// full column positions (getJacobian / getJtJR parity).
// alignRowGroups: every multi-row unit starts on a row that is a multiple of 4 (the rows in between stay zero), so that four
// consecutive rows of a column are one aligned 16-byte piece of the K-major device Jacobian (tile-sparse Gram kernel).
// columnOrder (optional, compact only): model parameters in the order the device columns should take (a permutation
// of the enabled parameters, e.g. the Cholesky elimination order); default = ascending enabled parameters.
std::string buildPlan(const HostCharacter& ch, const std::vector<HostErrorFunction>& efs, const std::vector<uint8_t>& enabled, bool compact, Plan& out,
                      const std::vector<int32_t>* columnOrder = nullptr, bool alignRowGroups = false);

// getJacobianSize() of a block (joint_error_function-inl.h:300-302, state_error_function.cpp:394-404,
// limit_error_function.cpp:1138-1161)
int32_t jacobianBlockSize(const HostCharacter& ch, const HostErrorFunction& ef);

} // namespace mb2

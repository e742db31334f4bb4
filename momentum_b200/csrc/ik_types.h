// Shared host/device data model of the batched Gauss-Newton IK path.
//
// Vocabulary follows the reference (momentum/): a *character* is Skeleton + ParameterTransform +
// ParameterLimits; a *solver function* is a batch of SkeletonSolverFunctionT<float> sharing one
// character and one constraint topology; error functions contribute *units* (one constraint /
// state joint / limit = one group of residual rows) and *cells* (one (unit, model-parameter) block
// of the Jacobian, with the chain-rule contributions that the reference's ancestor walk would add
// into it, joint_error_function-inl.h:229-294).
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define MB2_HD __host__ __device__ __forceinline__
#else
#define MB2_HD inline
#endif

namespace mb2 {

constexpr int kParametersPerJoint = 7; // character/types.h:21
constexpr int kJointStateStride = 17;  // t(3) q(4) s(1) rotationAxis(9); odd => bank-conflict-free per-joint access
constexpr float kLn2 = 0.69314718055994530942f; // math/constants.h:40
constexpr float kPi = 3.14159265358979323846f;

enum UnitKind : int32_t {
  kUnitPosition = 0,        // position_error_function.cpp:15-27
  kUnitOrientation = 1,     // orientation_error_function.cpp:15-40
  kUnitOrientationRotDiff = 2, // orientation_error_function.cpp:43-65
  kUnitStateMatrix = 3,     // state_error_function.cpp:407-558, RotationMatrixDifference
  kUnitStateLogMap = 4,     // ... QuaternionLogMap
  kUnitLimitMinMax = 5,     // limit_error_function.cpp:459-503
  kUnitLimitMinMaxJoint = 6, // :505-558
  kUnitLimitLinear = 7,     // :560-598
  kUnitLimitLinearJoint = 8, // :600-656
  kUnitLimitHalfPlane = 9,  // :658-699
  kUnitLimitEllipsoid = 10, // :701-785
  kUnitPlane = 11,          // plane_error_function.cpp:49-70 (one row; half-plane mode clamps)
  kUnitModelParameter = 12, // model_parameters_error_function.cpp:90-133 (one row per enabled parameter with target weight > 0)
};

enum LossType : int32_t { kLossL2 = 0, kLossL1 = 1, kLossCauchy = 2, kLossWelsch = 3, kLossGeneral = 4 };

// One error function block (SkeletonErrorFunctionT::weight_, GeneralizedLossT, StateErrorFunction weights)
struct EfDesc {
  float weight;
  int32_t lossType;
  float alpha;
  float invC2;
  float posWgt, rotWgt; // StateErrorFunctionT::posWgt_/rotWgt_
  int32_t kind;         // 0 pos, 1 ori, 2 rotdiff, 3 state, 4 limit, 5 plane, 6 model parameters
  int32_t halfPlane;    // PlaneErrorFunctionT(above)
};

struct UnitDesc {
  int32_t kind;      // UnitKind
  int32_t ef;        // index into EfDesc table
  int32_t joint;     // parent joint (pos/ori/ellipsoid), state joint
  int32_t row0;      // first residual row of this unit
  int32_t numRows;   // 3 / 9 / 12 / 6 / 1 / 3
  int32_t targetOff; // float offset into the per-instance target record (-1: none)
  int32_t weightIdx; // index into constraint-weight array (-1: none)
  int32_t recOff;    // float offset of this unit's evaluation record in the per-instance scratch
  int32_t i[4];      // limit indices (model parameter / joint-parameter rows), ellipsoidParent in i[0]
  float f[8];        // offset (3 or 4) | posW, rotW | limit floats f0..f3
  int32_t extra;     // float offset into limitData (ellipsoid matrices), -1 none
  int32_t pad[3];     // [0] limit gated off by the enabled set (zero rows); [1] error-only unit (ModelParameters with a negative target weight); [2] Position: per-instance offset stored after the target
};

// One Jacobian cell: rows [unit.row0, +numRows) x column `col`
struct CellDesc {
  uint16_t unit;
  uint16_t col;           // device column of J (model parameter index, or its rank in the enabled list when compacted)
  uint32_t contribBegin;  // into ContribDesc table
  uint16_t contribCount;  // 0 for limit cells that only scale by coef
  uint16_t quadStride;    // strip layout: strips between consecutive row quads of this cell's unit (same tile column)
  float coef;             // static coefficient for limit cells
  uint32_t stripOff;      // strip layout: float offset of (first row quad, this column) in the instance's strip buffer
};

// One chain-rule contribution: joint-parameter (joint, dof) with ParameterTransform coefficient
struct ContribDesc {
  uint16_t joint;
  uint16_t dof; // 0-2 translation, 3-5 rotation, 6 scale
  float coef;
};

// Everything the FK / residual / Jacobian kernels need (device pointers), passed by value.
struct FunctionTables {
  // character
  int32_t numJoints, numParams;
  const int32_t* parent;     // [J]
  const float* offset;       // [J*3]
  const float* prerot;       // [J*4] xyzw
  const int32_t* ptOuter;    // [7J+1]
  const int32_t* ptInner;    // [nnz]
  const float* ptVals;       // [nnz]
  const float* ptOffsets;    // [7J]
  int32_t numLevels;         // depth levels of the joint tree
  const int32_t* levelStart; // [numLevels+1]
  const int32_t* levelJoints; // [J] joints sorted by depth
  // objective
  int32_t numEf, numUnits, numCells;
  const EfDesc* efs;
  const UnitDesc* units;
  const CellDesc* cells;
  const ContribDesc* contribs;
  const float* limitData;
  int32_t targetStride; // floats per instance
  int32_t recStride;    // floats per instance of evaluation records
  int32_t numRows;      // active residual rows m (unpadded)
  int32_t ldJ;          // row stride of a Jacobian column (m padded to 32)
  int32_t numCols;      // Jacobian columns held on the device (all n, or only the enabled ones when compacted);
                        // the device matrix has numCols + 1 columns: the last one is the residual vector
  int32_t weightsPerInstance; // 0: cweights [numWeights] shared, 1: [B][numWeights]
  int32_t numWeights;
  // table sizes (the sweep kernel stages every table in shared memory when they fit)
  int32_t ptNnz, numContribs, numLimitData;
  // Jacobian output layout. 0: K-major matrix [numCols + 1][ldJ] (column numCols = residual). 1: strips (GramPlan in
  // ik_chol_sched.h): [numStrips][16 columns][4 rows] then the residual at residOff, rows numbered with 4-aligned row groups.
  int32_t stripMode, residOff;
  size_t jacobianStride;     // floats per instance in the Jacobian buffer
};

} // namespace mb2

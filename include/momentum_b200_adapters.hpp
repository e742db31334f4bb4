// Header-only C++ host side above the C-ABI (include/momentum_b200.h).
//
// Part 1 (always available, C++17, no third-party types): RAII mirrors of the reference classes on the
// hot path, batched over B independent IK instances:
//     momentum_b200::Character                     <- momentum::Character (skeleton + parameterTransform + parameterLimits)
//     momentum_b200::BatchedSkeletonSolverFunction <- momentum::SkeletonSolverFunctionT<float>  (skeleton_solver_function.h:21-95)
//     momentum_b200::BatchedGaussNewtonSolver      <- momentum::GaussNewtonSolverT<float>       (gauss_newton_solver.h:67-137)
//     momentum_b200::GaussNewtonSolverOptions      <- momentum::GaussNewtonSolverOptions        (gauss_newton_solver.h:17-59)
// Errors are rethrown as std::runtime_error like MT_CHECK / MT_THROW (common/exception.h:31,60-67).
//
// Part 2 (compiled only when momentum's headers are on the include path):
//     momentum_b200::makeCharacter(const momentum::Character&)       translates skeleton / parameterTransform / limits
//     momentum_b200::CudaSkeletonSolverFunction : momentum::SolverFunctionT<float>   (single instance; getError / getJacobian / getJtJR)
// so that existing callers can hand the solver function to momentum's own solvers; the batched device-side Gauss-Newton loop is
// reached through BatchedGaussNewtonSolver of part 1 (wrapping it in a momentum::SolverT subclass is a few lines on top of it).
// Part 2 cannot be compiled in the development image (Eigen 5, ms-gsl, fmt are absent); see INTEGRATION.md.
#pragma once

#include <bitset>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "momentum_b200.h"

namespace momentum_b200 {

inline void check(int rc) {
  if (rc != MB2_OK) throw std::runtime_error(mb2_last_error());
}

using ParameterSet = std::bitset<MB2_MAX_MODEL_PARAMETERS>; // math/types.h:426-429

inline void toWords(const ParameterSet& ps, uint64_t words[MB2_PARAMETER_SET_WORDS]) {
  for (int w = 0; w < MB2_PARAMETER_SET_WORDS; ++w) words[w] = 0;
  for (size_t i = 0; i < ps.size(); ++i)
    if (ps.test(i)) words[i >> 6] |= (uint64_t(1) << (i & 63));
}

struct GaussNewtonSolverOptions { // solver.h:19-34 + gauss_newton_solver.h:17-59
  size_t minIterations = 1;
  size_t maxIterations = 2;
  float threshold = 1.0f;
  bool verbose = false;
  float regularization = 0.05f;
  bool doLineSearch = false;
  bool useBlockJtJ = false;
  size_t targetRowsPerChunk = SIZE_MAX;
  // device extensions
  bool subsetLineSearch = false;
  mb2_jtj_mode jtjMode = MB2_JTJ_AUTO;
  mb2_cholesky_mode choleskyMode = MB2_CHOLESKY_AUTO;
  mb2_fused_mode fusedMode = MB2_FUSED_AUTO;
  bool storeErrorHistory = false;

  mb2_gauss_newton_options c() const {
    mb2_gauss_newton_options o;
    mb2_default_gauss_newton_options(&o);
    o.min_iterations = minIterations;
    o.max_iterations = maxIterations;
    o.threshold = threshold;
    o.verbose = verbose;
    o.regularization = regularization;
    o.do_line_search = doLineSearch;
    o.use_block_jtj = useBlockJtJ;
    o.target_rows_per_chunk = targetRowsPerChunk;
    o.subset_line_search = subsetLineSearch;
    o.jtj_mode = jtjMode;
    o.cholesky_mode = choleskyMode;
    o.fused_mode = fusedMode;
    o.store_error_history = storeErrorHistory;
    return o;
  }
};

class Character {
 public:
  Character(int device, const std::vector<int32_t>& parents, const std::vector<float>& translationOffsets /*3J*/,
            const std::vector<float>& preRotations /*4J xyzw*/, int32_t numModelParameters, const std::vector<int32_t>& transformOuter /*7J+1*/,
            const std::vector<int32_t>& transformInner, const std::vector<float>& transformValues, const std::vector<float>& transformOffsets /*7J*/) {
    check(mb2_character_create(device, int32_t(parents.size()), parents.data(), translationOffsets.data(), preRotations.data(), numModelParameters,
                               transformOuter.data(), transformInner.data(), transformValues.data(), transformOffsets.data(), &h_));
  }
  ~Character() { mb2_character_destroy(h_); }
  Character(const Character&) = delete;
  Character& operator=(const Character&) = delete;
  void setParameterLimits(const std::vector<mb2_parameter_limit>& limits) {
    check(mb2_character_set_parameter_limits(h_, int32_t(limits.size()), limits.data()));
  }
  mb2_character* handle() const { return h_; }

 private:
  mb2_character* h_{nullptr};
};

class BatchedSkeletonSolverFunction {
 public:
  BatchedSkeletonSolverFunction(const Character& character, int32_t batch) { check(mb2_solver_function_create(character.handle(), batch, &h_)); }
  ~BatchedSkeletonSolverFunction() { mb2_solver_function_destroy(h_); }
  BatchedSkeletonSolverFunction(const BatchedSkeletonSolverFunction&) = delete;
  BatchedSkeletonSolverFunction& operator=(const BatchedSkeletonSolverFunction&) = delete;

  [[nodiscard]] size_t getNumParameters() const { return size_t(mb2_solver_function_num_parameters(h_)); }
  [[nodiscard]] size_t getActualParameters() const { return size_t(mb2_solver_function_actual_parameters(h_)); }
  [[nodiscard]] int32_t batch() const { return mb2_solver_function_batch(h_); }

  // addErrorFunction(std::make_shared<PositionErrorFunction>(...)) + setConstraints
  int addPositionErrorFunction(float weight, const std::vector<int32_t>& parents, const std::vector<float>& offsets, const std::vector<float>& weights,
                               float lossAlpha = 2.f, float lossC = 1.f) {
    int32_t idx = -1;
    check(mb2_add_position_error_function(h_, weight, lossAlpha, lossC, int32_t(parents.size()), parents.data(), offsets.data(), weights.data(), &idx));
    return idx;
  }
  // PlaneErrorFunctionT(character, above) + setConstraints; targets [B][nc*4] = normal, d per instance
  int addPlaneErrorFunction(float weight, const std::vector<int32_t>& parents, const std::vector<float>& offsets, const std::vector<float>& weights,
                            bool above = false, float lossAlpha = 2.f, float lossC = 1.f) {
    int32_t idx = -1;
    check(mb2_add_plane_error_function(h_, weight, lossAlpha, lossC, above, int32_t(parents.size()), parents.data(), offsets.data(), weights.data(), &idx));
    return idx;
  }
  // ModelParametersErrorFunctionT + setTargetParameters (weights here, per-instance target parameters through setTargets)
  int addModelParametersErrorFunction(float weight, const std::vector<float>& targetWeights) {
    int32_t idx = -1;
    check(mb2_add_model_parameters_error_function(h_, weight, targetWeights.data(), &idx));
    return idx;
  }
  int addOrientationErrorFunction(float weight, const std::vector<int32_t>& parents, const std::vector<float>& offsetsXYZW, const std::vector<float>& weights,
                                  bool rotDiff = false, float lossAlpha = 2.f, float lossC = 1.f) {
    int32_t idx = -1;
    check(mb2_add_orientation_error_function(h_, weight, lossAlpha, lossC, rotDiff, int32_t(parents.size()), parents.data(), offsetsXYZW.data(), weights.data(),
                                             &idx));
    return idx;
  }
  int addStateErrorFunction(float weight, mb2_rotation_error_type type, float posWgt, float rotWgt, const std::vector<float>& positionWeights,
                            const std::vector<float>& rotationWeights) {
    int32_t idx = -1;
    check(mb2_add_state_error_function(h_, weight, type, posWgt, rotWgt, positionWeights.data(), rotationWeights.data(), &idx));
    return idx;
  }
  int addLimitErrorFunction(float weight, float lossAlpha = 2.f, float lossC = 1.f) {
    int32_t idx = -1;
    check(mb2_add_limit_error_function(h_, weight, lossAlpha, lossC, &idx));
    return idx;
  }
  void setTargets(int index, const std::vector<float>& targets) { check(mb2_set_targets(h_, index, targets.data())); }
  void setEnabledParameters(const ParameterSet& ps) {
    uint64_t w[MB2_PARAMETER_SET_WORDS];
    toWords(ps, w);
    check(mb2_solver_function_set_enabled_parameters(h_, w));
  }
  // getError for every instance (skeleton_solver_function.cpp:64-83)
  std::vector<double> getError(const std::vector<float>& parameters) {
    std::vector<double> e(static_cast<size_t>(batch()), 0.0);
    check(mb2_solver_function_get_error(h_, parameters.data(), e.data()));
    return e;
  }
  mb2_solver_function* handle() const { return h_; }

 private:
  mb2_solver_function* h_{nullptr};
};

struct BatchedSolveResult {
  std::vector<double> errors;      // what SolverT::solve returns, per instance
  std::vector<int32_t> iterations; // doIteration calls, per instance
  std::vector<int32_t> status;     // mb2_instance_status
};

class BatchedGaussNewtonSolver {
 public:
  BatchedGaussNewtonSolver(const GaussNewtonSolverOptions& options, BatchedSkeletonSolverFunction* function) : fn_(function) {
    const mb2_gauss_newton_options o = options.c();
    check(mb2_solver_create(function->handle(), &o, &h_));
  }
  ~BatchedGaussNewtonSolver() { mb2_solver_destroy(h_); }
  BatchedGaussNewtonSolver(const BatchedGaussNewtonSolver&) = delete;
  BatchedGaussNewtonSolver& operator=(const BatchedGaussNewtonSolver&) = delete;

  [[nodiscard]] std::string_view getName() const { return "GaussNewton"; }
  void setOptions(const GaussNewtonSolverOptions& options) {
    const mb2_gauss_newton_options o = options.c();
    check(mb2_solver_set_options(h_, &o));
  }
  void setEnabledParameters(const ParameterSet& ps) {
    uint64_t w[MB2_PARAMETER_SET_WORDS];
    toWords(ps, w);
    check(mb2_solver_set_enabled_parameters(h_, w));
  }
  // SolverT::solve for the whole batch: parameters [B * n] in/out
  BatchedSolveResult solve(std::vector<float>& parameters) {
    const size_t B = size_t(fn_->batch());
    if (parameters.size() != B * fn_->getNumParameters()) throw std::runtime_error("parameters size must be batch * numParameters"); // solver.cpp:77
    BatchedSolveResult r;
    r.errors.resize(B);
    r.iterations.resize(B);
    r.status.resize(B);
    check(mb2_solver_solve(h_, parameters.data(), r.errors.data(), r.iterations.data(), r.status.data()));
    return r;
  }

 private:
  BatchedSkeletonSolverFunction* fn_;
  mb2_solver* h_{nullptr};
};

} // namespace momentum_b200

// ------------------------------------------------------------------------------------------------------------------
// Part 2: drop-in subclasses of momentum's own interfaces (needs momentum + Eigen headers)
// ------------------------------------------------------------------------------------------------------------------
#if defined(__has_include)
#if __has_include(<momentum/solver/solver.h>) && __has_include(<momentum/character_solver/skeleton_solver_function.h>)
#define MOMENTUM_B200_HAVE_MOMENTUM 1
#include <momentum/character/character.h>
#include <momentum/character_solver/limit_error_function.h>
#include <momentum/character_solver/orientation_error_function.h>
#include <momentum/character_solver/position_error_function.h>
#include <momentum/character_solver/state_error_function.h>
#include <momentum/solver/solver.h>
#include <momentum/solver/solver_function.h>

namespace momentum_b200 {

inline std::unique_ptr<Character> makeCharacter(int device, const momentum::Character& c) {
  const auto& sk = c.skeleton;
  const auto& pt = c.parameterTransform;
  const size_t J = sk.joints.size();
  std::vector<int32_t> parents(J);
  std::vector<float> off(3 * J), pre(4 * J);
  for (size_t j = 0; j < J; ++j) {
    parents[j] = sk.joints[j].parent == momentum::kInvalidIndex ? -1 : int32_t(sk.joints[j].parent);
    for (int k = 0; k < 3; ++k) off[3 * j + k] = sk.joints[j].translationOffset[k];
    const auto& q = sk.joints[j].preRotation.coeffs(); // x,y,z,w
    for (int k = 0; k < 4; ++k) pre[4 * j + k] = q[k];
  }
  const auto& T = pt.transform; // SparseRowMatrix<float>, 7J x n
  std::vector<int32_t> outer(T.outerIndexPtr(), T.outerIndexPtr() + T.rows() + 1), inner(T.innerIndexPtr(), T.innerIndexPtr() + T.nonZeros());
  std::vector<float> vals(T.valuePtr(), T.valuePtr() + T.nonZeros()), offs(pt.offsets.data(), pt.offsets.data() + pt.offsets.size());
  auto out = std::make_unique<Character>(device, parents, off, pre, int32_t(T.cols()), outer, inner, vals, offs);
  std::vector<mb2_parameter_limit> lim;
  for (const auto& l : c.parameterLimits) {
    mb2_parameter_limit m{};
    m.type = int32_t(l.type);
    m.weight = l.weight;
    switch (l.type) {
      case momentum::MinMax: m.i[0] = int32_t(l.data.minMax.parameterIndex); m.f[0] = l.data.minMax.limits[0]; m.f[1] = l.data.minMax.limits[1]; break;
      case momentum::MinMaxJoint:
      case momentum::MinMaxJointPassive:
        m.i[0] = int32_t(l.data.minMaxJoint.jointIndex); m.i[1] = int32_t(l.data.minMaxJoint.jointParameter);
        m.f[0] = l.data.minMaxJoint.limits[0]; m.f[1] = l.data.minMaxJoint.limits[1]; break;
      case momentum::Linear:
        m.i[0] = int32_t(l.data.linear.referenceIndex); m.i[1] = int32_t(l.data.linear.targetIndex);
        m.f[0] = l.data.linear.scale; m.f[1] = l.data.linear.offset; m.f[2] = l.data.linear.rangeMin; m.f[3] = l.data.linear.rangeMax; break;
      case momentum::LinearJoint:
        m.i[0] = int32_t(l.data.linearJoint.referenceJointIndex); m.i[1] = int32_t(l.data.linearJoint.referenceJointParameter);
        m.i[2] = int32_t(l.data.linearJoint.targetJointIndex); m.i[3] = int32_t(l.data.linearJoint.targetJointParameter);
        m.f[0] = l.data.linearJoint.scale; m.f[1] = l.data.linearJoint.offset; m.f[2] = l.data.linearJoint.rangeMin; m.f[3] = l.data.linearJoint.rangeMax; break;
      case momentum::HalfPlane:
        m.i[0] = int32_t(l.data.halfPlane.param1); m.i[1] = int32_t(l.data.halfPlane.param2);
        m.f[0] = l.data.halfPlane.normal[0]; m.f[1] = l.data.halfPlane.normal[1]; m.f[2] = l.data.halfPlane.offset; break;
      case momentum::Ellipsoid: {
        m.i[0] = int32_t(l.data.ellipsoid.ellipsoidParent); m.i[1] = int32_t(l.data.ellipsoid.parent);
        const auto& E = l.data.ellipsoid.ellipsoid.matrix();
        const auto& Ei = l.data.ellipsoid.ellipsoidInv.matrix();
        for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 4; ++cc) { m.f[4 * r + cc] = E(r, cc); m.f[12 + 4 * r + cc] = Ei(r, cc); }
        for (int k = 0; k < 3; ++k) m.f[24 + k] = l.data.ellipsoid.offset[k];
        break;
      }
      default: throw std::runtime_error("Unknown parameter type for joint limit");
    }
    lim.push_back(m);
  }
  out->setParameterLimits(lim);
  return out;
}

// A SolverFunctionT<float> whose getJtJR / getError / updateParameters run on the GPU for ONE instance
// (batch = 1); the stock momentum::GaussNewtonSolverT with useBlockJtJ = true then works unmodified
// (it only calls getJtJR, updateParameters and getError: gauss_newton_solver.cpp:75,286,305).
// For throughput use BatchedGaussNewtonSolver: one launch sequence for thousands of instances.
class CudaSkeletonSolverFunction : public momentum::SolverFunctionT<float> {
 public:
  CudaSkeletonSolverFunction(const momentum::Character& character, int device = 0)
      : character_(makeCharacter(device, character)), fn_(std::make_unique<BatchedSkeletonSolverFunction>(*character_, 1)) {
    this->numParameters_ = fn_->getNumParameters();
    this->actualParameters_ = this->numParameters_;
  }
  BatchedSkeletonSolverFunction& batched() { return *fn_; }

  double getError(const momentum::VectorX<float>& parameters) final {
    double e = 0;
    check(mb2_solver_function_get_error(fn_->handle(), parameters.data(), &e));
    return e;
  }
  double getGradient(const momentum::VectorX<float>& parameters, momentum::VectorX<float>& gradient) final {
    momentum::MatrixX<float> jtj;
    const double e = getJtJR(parameters, jtj, gradient);
    gradient *= 2.0f; // gradient = 2 J^T r (error_function_helpers.cpp:245-259)
    return e;
  }
  double getJtJR(const momentum::VectorX<float>& parameters, momentum::MatrixX<float>& jtj, momentum::VectorX<float>& jtr) final {
    const Eigen::Index ap = Eigen::Index(fn_->getActualParameters());
    Eigen::Matrix<float, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> h(ap, ap);
    h.setZero();
    jtr.setZero(ap);
    double e = 0;
    check(mb2_solver_function_get_jtjr(fn_->handle(), parameters.data(), MB2_JTJ_AUTO, h.data(), jtr.data(), &e));
    jtj = h; // lower triangle valid, as after selfadjointView<Lower>().rankUpdate
    return e;
  }
  void initializeJacobianComputation(const momentum::VectorX<float>& parameters) final { lastParameters_ = parameters; }
  [[nodiscard]] size_t getJacobianBlockCount() const final { return 1; }
  [[nodiscard]] size_t getJacobianBlockSize(size_t) const final { return size_t(mb2_solver_function_jacobian_rows(fn_->handle())); }
  double computeJacobianBlock(const momentum::VectorX<float>& parameters, size_t, Eigen::Ref<momentum::MatrixX<float>> jacobianBlock,
                              Eigen::Ref<momentum::VectorX<float>> residualBlock, size_t& actualRows) final {
    const int rows = mb2_solver_function_jacobian_rows(fn_->handle());
    momentum::MatrixX<float> J(rows, Eigen::Index(this->numParameters_));
    momentum::VectorX<float> r(rows);
    double e = 0;
    int32_t ar = 0;
    check(mb2_solver_function_get_jacobian(fn_->handle(), parameters.data(), J.data(), r.data(), &e, &ar));
    jacobianBlock.topRows(rows) = J;
    residualBlock.head(rows) = r;
    actualRows = size_t(ar);
    return e;
  }
  void updateParameters(momentum::VectorX<float>& parameters, const momentum::VectorX<float>& delta) final { parameters -= delta; }
  void setEnabledParameters(const momentum::ParameterSet& ps) final {
    fn_->setEnabledParameters(ps);
    this->actualParameters_ = fn_->getActualParameters();
  }

 private:
  std::unique_ptr<Character> character_;
  std::unique_ptr<BatchedSkeletonSolverFunction> fn_;
  momentum::VectorX<float> lastParameters_;
};

} // namespace momentum_b200
#endif
#endif

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
// Part 2 (compiled only when momentum's headers are on the include path, C++20):
//     momentum_b200::makeCharacter(const momentum::Character&)       translates skeleton / parameterTransform / limits
//     momentum_b200::CudaSkeletonSolverFunction : momentum::SolverFunctionT<float>   same ctor shape as SkeletonSolverFunctionT,
//                                                                    addErrorFunction(shared_ptr<SkeletonErrorFunctionT<float>>)
//     momentum_b200::CudaGaussNewtonSolver : momentum::SolverT<float>
// momentum + Eigen 5 are not in the development image; tests/mock_momentum/ holds signature-level stand-ins of the handful of
// momentum headers these classes touch, and tests/test_cpp_adapters.py compiles and RUNS part 2 against them.
#pragma once

#include <algorithm>
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
  mb2_linear_solver linearSolver = MB2_LINEAR_SOLVER_CHOLESKY; // MB2_LINEAR_SOLVER_QR: GaussNewtonSolverQRT's step; MB2_LINEAR_SOLVER_TRUST_REGION_QR: TrustRegionQRT
  float trustRegionRadius = 1.0f;                              // TrustRegionQROptions::trustRegionRadius_ (trust_region_qr.h:23)
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
    o.linear_solver = linearSolver;
    o.trust_region_radius = trustRegionRadius;
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
  // ConstraintData::weight of a Position / Orientation / Plane block: [nc] shared by the batch, or [B * nc] per instance
  void setConstraintWeights(int index, const std::vector<float>& weights, bool perInstance = false) {
    check(mb2_set_constraint_weights(h_, index, weights.data(), perInstance ? 1 : 0));
  }
  void setErrorFunctionWeight(int index, float weight) { check(mb2_set_error_function_weight(h_, index, weight)); } // SkeletonErrorFunctionT::setWeight
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
  BatchedGaussNewtonSolver(const GaussNewtonSolverOptions& options, BatchedSkeletonSolverFunction* function) : fn_(function), historyStride_(std::max<size_t>(options.maxIterations, 1)) {
    const mb2_gauss_newton_options o = options.c();
    check(mb2_solver_create(function->handle(), &o, &h_));
  }
  ~BatchedGaussNewtonSolver() { mb2_solver_destroy(h_); }
  BatchedGaussNewtonSolver(const BatchedGaussNewtonSolver&) = delete;
  BatchedGaussNewtonSolver& operator=(const BatchedGaussNewtonSolver&) = delete;

  [[nodiscard]] std::string_view getName() const { return "GaussNewton"; }
  void setOptions(const GaussNewtonSolverOptions& options) {
    const mb2_gauss_newton_options o = options.c();
    historyStride_ = std::max<size_t>(options.maxIterations, 1);
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
  // getErrorHistory (solver.h:90) of instance 0, first `count` iterations (needs storeErrorHistory)
  std::vector<double> getErrorHistory(size_t count, size_t maxIterations = 0) {
    (void)maxIterations;
    const size_t B = size_t(fn_->batch());
    std::vector<double> all(B * std::max<size_t>(historyStride_, 1), 0.0);
    check(mb2_solver_get_error_history(h_, all.data()));
    all.resize(std::min(count, historyStride_));
    return all;
  }

 private:
  BatchedSkeletonSolverFunction* fn_;
  size_t historyStride_{1};
  mb2_solver* h_{nullptr};
};

} // namespace momentum_b200

// ------------------------------------------------------------------------------------------------------------------
// Part 2: drop-in subclasses of momentum's own interfaces (compiled when momentum's headers are on the include path; C++20 like
// momentum itself). tests/mock_momentum/ carries signature-level stand-ins of those headers so that this part is compiled and run in CI.
//
//   CudaSkeletonSolverFunction : momentum::SolverFunctionT<float>   same constructor shape as SkeletonSolverFunctionT
//        (skeleton_solver_function.h:23-26); addErrorFunction(shared_ptr<SkeletonErrorFunctionT<float>>) translates Position /
//        Orientation / OrientationRotDiff / Plane / State / Limit / ModelParameters objects (dynamic_pointer_cast, then their public
//        getters) into device tables and throws std::runtime_error for anything else. momentum's stock GaussNewtonSolverT with
//        useBlockJtJ = true drives it unmodified (it only calls getJtJR, updateParameters, getError: gauss_newton_solver.cpp:75,286,305).
//   CudaGaussNewtonSolver : momentum::SolverT<float>                initializeSolver / doIteration / getName (solver.h:49,96,99): every
//        doIteration is one device Gauss-Newton iteration; solveOnDevice() runs the whole SolverT loop in one C-ABI call.
// Some of the state an adapter needs has no public getter upstream (loss parameters, PlaneErrorFunctionT::halfPlane_,
// StateErrorFunctionT::rotationErrorType_, LimitErrorFunctionT::limits_). It is read through explicit-instantiation member pointers
// (standard C++: access checking does not apply to the arguments of an explicit instantiation), so momentum needs no change.
// ------------------------------------------------------------------------------------------------------------------
#if defined(__has_include)
#if __has_include(<momentum/solver/solver.h>) && __has_include(<momentum/character_solver/skeleton_solver_function.h>)
#define MOMENTUM_B200_HAVE_MOMENTUM 1
#include <momentum/character/character.h>
#include <momentum/character_solver/limit_error_function.h>
#include <momentum/character_solver/model_parameters_error_function.h>
#include <momentum/character_solver/orientation_error_function.h>
#include <momentum/character_solver/plane_error_function.h>
#include <momentum/character_solver/position_error_function.h>
#include <momentum/character_solver/state_error_function.h>
#include <momentum/solver/gauss_newton_solver.h>
#include <momentum/solver/solver.h>
#include <momentum/solver/solver_function.h>

#include <cmath>
#include <span>

namespace momentum_b200 {

namespace detail {
template <class Tag, typename Tag::type Member>
struct Expose {
  friend constexpr typename Tag::type exposed(Tag) { return Member; }
};
#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wnon-template-friend"
#endif
#define MB2_EXPOSE(Tag, Class, MemberType, member) \
  struct Tag {                                      \
    using type = MemberType Class::*;               \
    friend constexpr type exposed(Tag);             \
  };                                                \
  template struct Expose<Tag, &Class::member>;
using PositionBase = momentum::JointErrorFunctionT<float, momentum::PositionDataT<float>>;
using OrientationBase = momentum::JointErrorFunctionT<float, momentum::OrientationDataT<float>, 9, 3, 0>;
using PlaneBase = momentum::JointErrorFunctionT<float, momentum::PlaneDataT<float>, 1>;
using Loss = momentum::GeneralizedLossT<float>;
MB2_EXPOSE(PositionLossTag, PositionBase, const Loss, loss_)
MB2_EXPOSE(OrientationLossTag, OrientationBase, const Loss, loss_)
MB2_EXPOSE(PlaneLossTag, PlaneBase, const Loss, loss_)
MB2_EXPOSE(LimitLossTag, momentum::LimitErrorFunctionT<float>, const Loss, loss_)
MB2_EXPOSE(LimitLimitsTag, momentum::LimitErrorFunctionT<float>, momentum::ParameterLimits, limits_)
MB2_EXPOSE(LossAlphaTag, Loss, const float, alpha_)
MB2_EXPOSE(PlaneHalfTag, momentum::PlaneErrorFunctionT<float>, bool, halfPlane_)
MB2_EXPOSE(StateRotTypeTag, momentum::StateErrorFunctionT<float>, const momentum::RotationErrorType, rotationErrorType_)
#undef MB2_EXPOSE
#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic pop
#endif
inline float lossAlpha(const Loss& l) { return l.*exposed(LossAlphaTag{}); }
inline float lossC(const Loss& l) { return 1.0f / std::sqrt(l.invC2()); }

inline std::vector<mb2_parameter_limit> translateLimits(const momentum::ParameterLimits& limits) {
  std::vector<mb2_parameter_limit> out;
  for (const auto& l : limits) {
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
        const auto E = l.data.ellipsoid.ellipsoid.matrix();
        const auto Ei = l.data.ellipsoid.ellipsoidInv.matrix();
        for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 4; ++cc) { m.f[4 * r + cc] = E(r, cc); m.f[12 + 4 * r + cc] = Ei(r, cc); }
        for (int k = 0; k < 3; ++k) m.f[24 + k] = l.data.ellipsoid.offset[k];
        break;
      }
      default: throw std::runtime_error("Unknown parameter type for joint limit");
    }
    out.push_back(m);
  }
  return out;
}
} // namespace detail

// momentum::Character (skeleton) + a ParameterTransformT<float> -> device character (limits are attached by the caller)
inline std::unique_ptr<Character> makeCharacter(int device, const momentum::Skeleton& sk, const momentum::ParameterTransformT<float>& pt) {
  const size_t J = sk.joints.size();
  std::vector<int32_t> parents(J);
  std::vector<float> off(3 * J), pre(4 * J);
  for (size_t j = 0; j < J; ++j) {
    parents[j] = sk.joints[j].parent == momentum::kInvalidIndex ? -1 : int32_t(sk.joints[j].parent);
    for (int k = 0; k < 3; ++k) off[3 * j + k] = sk.joints[j].translationOffset[k];
    const auto& q = sk.joints[j].preRotation;
    pre[4 * j] = q.x(); pre[4 * j + 1] = q.y(); pre[4 * j + 2] = q.z(); pre[4 * j + 3] = q.w();
  }
  const auto& T = pt.transform; // SparseRowMatrix<float>, 7J x n
  std::vector<int32_t> outer(T.outerIndexPtr(), T.outerIndexPtr() + T.rows() + 1), inner(T.innerIndexPtr(), T.innerIndexPtr() + T.nonZeros());
  std::vector<float> vals(T.valuePtr(), T.valuePtr() + T.nonZeros()), offs(pt.offsets.data(), pt.offsets.data() + pt.offsets.size());
  return std::make_unique<Character>(device, parents, off, pre, int32_t(T.cols()), outer, inner, vals, offs);
}
inline std::unique_ptr<Character> makeCharacter(int device, const momentum::Character& c) {
  auto out = makeCharacter(device, c.skeleton, c.parameterTransform);
  out->setParameterLimits(detail::translateLimits(c.parameterLimits));
  return out;
}

// One instance of SkeletonSolverFunctionT<float> evaluated on the GPU (batch of one through the batched C-ABI). For throughput use
// BatchedSkeletonSolverFunction / BatchedGaussNewtonSolver of part 1: one launch sequence for thousands of instances.
class CudaSkeletonSolverFunction : public momentum::SolverFunctionT<float> {
 public:
  using ErrorFunctionPtr = std::shared_ptr<momentum::SkeletonErrorFunctionT<float>>;

  CudaSkeletonSolverFunction(const momentum::Character& character, const momentum::ParameterTransformT<float>& parameterTransform,
                             std::span<const ErrorFunctionPtr> errorFunctions = {}, int device = 0)
      : character_(character), parameterTransform_(parameterTransform), device_(device) {
    this->numParameters_ = size_t(parameterTransform.transform.cols());
    this->actualParameters_ = this->numParameters_;
    enabled_.set();
    for (const auto& ef : errorFunctions) addErrorFunction(ef);
  }
  CudaSkeletonSolverFunction(const CudaSkeletonSolverFunction&) = delete;
  CudaSkeletonSolverFunction& operator=(const CudaSkeletonSolverFunction&) = delete;

  // skeleton_solver_function.cpp:161-169; the object is kept alive and re-read before every evaluation (setConstraints / setWeight /
  // setTargetState on it take effect like they do in momentum)
  void addErrorFunction(ErrorFunctionPtr ef) {
    if (!ef) throw std::runtime_error("addErrorFunction: null error function");
    (void)translate(*ef); // throws for unsupported classes now rather than at the first solve
    errorFunctions_.push_back(std::move(ef));
  }
  void clearErrorFunctions() { errorFunctions_.clear(); }
  [[nodiscard]] const std::vector<ErrorFunctionPtr>& getErrorFunctions() const { return errorFunctions_; }
  [[nodiscard]] const momentum::Character& getCharacter() const { return character_; }
  [[nodiscard]] const momentum::ParameterTransformT<float>* getParameterTransform() const { return &parameterTransform_; }

  double getError(const momentum::VectorX<float>& parameters) final {
    sync();
    double e = 0;
    check(mb2_solver_function_get_error(fn_->handle(), parameters.data(), &e));
    return e;
  }
  // gradient = 2 J^T r over all parameters (skeleton_solver_function.cpp:86-111 resizes it to parameters.size())
  double getGradient(const momentum::VectorX<float>& parameters, momentum::VectorX<float>& gradient) final {
    momentum::MatrixX<float> jtj;
    momentum::VectorX<float> jtr;
    const double e = getJtJR(parameters, jtj, jtr);
    gradient.setZero(parameters.size());
    for (Eigen::Index i = 0; i < jtr.size(); ++i) gradient(i) = 2.0f * jtr(i);
    return e;
  }
  double getJtJR(const momentum::VectorX<float>& parameters, momentum::MatrixX<float>& jtj, momentum::VectorX<float>& jtr) final {
    sync();
    const Eigen::Index ap = Eigen::Index(fn_->getActualParameters());
    std::vector<float> h(size_t(ap) * size_t(ap), 0.f), g(size_t(ap), 0.f);
    double e = 0;
    check(mb2_solver_function_get_jtjr(fn_->handle(), parameters.data(), MB2_JTJ_AUTO, h.data(), g.data(), &e));
    jtj.setZero(ap, ap); // lower triangle valid, as after selfadjointView<Lower>().rankUpdate (solver_function.cpp:113)
    jtr.setZero(ap);
    for (Eigen::Index i = 0; i < ap; ++i) {
      jtr(i) = g[size_t(i)];
      for (Eigen::Index j = 0; j <= i; ++j) jtj(i, j) = h[size_t(i) * size_t(ap) + size_t(j)];
    }
    return e;
  }
  void initializeJacobianComputation(const momentum::VectorX<float>&) final { sync(); }
  [[nodiscard]] size_t getJacobianBlockCount() const final { return 1; }
  [[nodiscard]] size_t getJacobianBlockSize(size_t) const final {
    const_cast<CudaSkeletonSolverFunction*>(this)->sync();
    return size_t(mb2_solver_function_jacobian_rows(fn_->handle()));
  }
  double computeJacobianBlock(const momentum::VectorX<float>& parameters, size_t, Eigen::Ref<momentum::MatrixX<float>> jacobianBlock,
                              Eigen::Ref<momentum::VectorX<float>> residualBlock, size_t& actualRows) final {
    sync();
    const int rows = mb2_solver_function_jacobian_rows(fn_->handle());
    const size_t n = this->numParameters_;
    std::vector<float> J(size_t(rows) * n, 0.f), r(size_t(rows), 0.f); // [n][rows]: column-major rows x n
    double e = 0;
    int32_t ar = 0;
    check(mb2_solver_function_get_jacobian(fn_->handle(), parameters.data(), J.data(), r.data(), &e, &ar));
    for (size_t c = 0; c < n; ++c)
      for (int k = 0; k < rows; ++k) jacobianBlock(k, Eigen::Index(c)) = J[c * size_t(rows) + size_t(k)];
    for (int k = 0; k < rows; ++k) residualBlock(k) = r[size_t(k)];
    actualRows = size_t(ar);
    return e;
  }
  void updateParameters(momentum::VectorX<float>& parameters, const momentum::VectorX<float>& delta) final { parameters -= delta; } // :153-159
  void setEnabledParameters(const momentum::ParameterSet& ps) final { // :45-61
    enabled_ = ps;
    size_t ap = 0;
    for (size_t i = 0; i < this->numParameters_; ++i) if (ps.test(i)) ap = i + 1;
    this->actualParameters_ = ap;
    for (auto& ef : errorFunctions_) ef->setEnabledParameters(ps);
    if (fn_) fn_->setEnabledParameters(ps);
  }

  // The device function of the current error-function set, in sync with the momentum objects
  BatchedSkeletonSolverFunction& batched() { sync(); return *fn_; }

  // Re-reads every error function: a change of topology (number / kind of blocks, parents, offsets, weights, loss, limits) rebuilds the
  // device function, anything else (targets, block weights) is a small upload.
  void sync() {
    std::vector<Block> blocks;
    blocks.reserve(errorFunctions_.size());
    const momentum::ParameterLimits* limits = nullptr;
    for (const auto& ef : errorFunctions_) {
      blocks.push_back(translate(*ef));
      if (blocks.back().limits) {
        if (limits != nullptr && limits != blocks.back().limits) throw std::runtime_error("CudaSkeletonSolverFunction supports one set of ParameterLimits per solver function");
        limits = blocks.back().limits;
      }
    }
    bool rebuild = !fn_ || blocks.size() != blocks_.size();
    for (size_t i = 0; !rebuild && i < blocks.size(); ++i) rebuild = !blocks[i].sameTopology(blocks_[i]);
    if (rebuild) {
      deviceCharacter_ = makeCharacter(device_, character_.skeleton, parameterTransform_);
      deviceCharacter_->setParameterLimits(detail::translateLimits(limits ? *limits : character_.parameterLimits));
      fn_ = std::make_unique<BatchedSkeletonSolverFunction>(*deviceCharacter_, 1);
      for (const Block& b : blocks) addBlock(b);
      fn_->setEnabledParameters(enabled_);
    }
    for (size_t i = 0; i < blocks.size(); ++i) {
      const Block& b = blocks[i];
      if (rebuild || b.weight != blocks_[i].weight) check(mb2_set_error_function_weight(fn_->handle(), int32_t(i), b.weight));
      if (!b.targets.empty() && (rebuild || b.targets != blocks_[i].targets)) fn_->setTargets(int(i), b.targets);
    }
    blocks_ = std::move(blocks);
  }

 private:
  struct Block {
    int kind = -1; // 0 position, 1 orientation, 2 rot-diff, 3 state, 4 limit, 5 plane, 6 model parameters
    float weight = 1.f, alpha = 2.f, c = 1.f, posWgt = 1.f, rotWgt = 1.f;
    bool above = false;
    int rotationErrorType = 0;
    std::vector<int32_t> parents;
    std::vector<float> offsets, cweights, jointPosW, jointRotW, paramWeights, targets;
    const momentum::ParameterLimits* limits = nullptr;
    [[nodiscard]] bool sameTopology(const Block& o) const {
      return kind == o.kind && alpha == o.alpha && c == o.c && posWgt == o.posWgt && rotWgt == o.rotWgt && above == o.above && rotationErrorType == o.rotationErrorType &&
             parents == o.parents && offsets == o.offsets && cweights == o.cweights && jointPosW == o.jointPosW && jointRotW == o.jointRotW &&
             paramWeights == o.paramWeights && limits == o.limits && (limits == nullptr || limitCount == o.limitCount);
    }
    size_t limitCount = 0;
  };

  template <class EF, class LossTag>
  static void jointCommon(const EF& ef, LossTag tag, Block& b) {
    const detail::Loss& loss = ef.*exposed(tag);
    b.alpha = detail::lossAlpha(loss);
    b.c = detail::lossC(loss);
    for (const auto& cst : ef.getConstraints()) {
      b.parents.push_back(int32_t(cst.parent));
      b.cweights.push_back(cst.weight);
    }
  }

  Block translate(const momentum::SkeletonErrorFunctionT<float>& ef) const {
    Block b;
    b.weight = ef.getWeight();
    if (const auto* p = dynamic_cast<const momentum::PositionErrorFunctionT<float>*>(&ef)) {
      b.kind = 0;
      jointCommon(*p, detail::PositionLossTag{}, b);
      for (const auto& cst : p->getConstraints())
        for (int k = 0; k < 3; ++k) { b.offsets.push_back(cst.offset[k]); b.targets.push_back(cst.target[k]); }
    } else if (const auto* o = dynamic_cast<const momentum::OrientationErrorFunctionT<float>*>(&ef)) {
      b.kind = 1;
      jointCommon(*o, detail::OrientationLossTag{}, b);
      quats(o->getConstraints(), b);
    } else if (const auto* rd = dynamic_cast<const momentum::OrientationRotDiffErrorFunctionT<float>*>(&ef)) {
      b.kind = 2;
      jointCommon(*rd, detail::OrientationLossTag{}, b);
      quats(rd->getConstraints(), b);
    } else if (const auto* pl = dynamic_cast<const momentum::PlaneErrorFunctionT<float>*>(&ef)) {
      b.kind = 5;
      jointCommon(*pl, detail::PlaneLossTag{}, b);
      b.above = pl->*exposed(detail::PlaneHalfTag{});
      for (const auto& cst : pl->getConstraints()) {
        for (int k = 0; k < 3; ++k) { b.offsets.push_back(cst.offset[k]); b.targets.push_back(cst.normal[k]); }
        b.targets.push_back(cst.d);
      }
    } else if (const auto* st = dynamic_cast<const momentum::StateErrorFunctionT<float>*>(&ef)) {
      b.kind = 3;
      b.posWgt = st->getPositionWeight();
      b.rotWgt = st->getRotationWeight();
      b.rotationErrorType = (st->*exposed(detail::StateRotTypeTag{})) == momentum::RotationErrorType::QuaternionLogMap ? 1 : 0;
      const size_t J = character_.skeleton.joints.size();
      const auto& pw = st->getPositionWeights();
      const auto& rw = st->getRotationWeights();
      const auto& tgt = st->getTargetState();
      for (size_t j = 0; j < J; ++j) {
        b.jointPosW.push_back(j < size_t(pw.size()) ? pw(Eigen::Index(j)) : 0.f);
        b.jointRotW.push_back(j < size_t(rw.size()) ? rw(Eigen::Index(j)) : 0.f);
        if (j < tgt.size()) {
          const auto& t = tgt[j];
          b.targets.insert(b.targets.end(), {t.translation[0], t.translation[1], t.translation[2], t.rotation.x(), t.rotation.y(), t.rotation.z(), t.rotation.w(), t.scale});
        } else {
          b.targets.insert(b.targets.end(), {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 1.f});
        }
      }
    } else if (const auto* lm = dynamic_cast<const momentum::LimitErrorFunctionT<float>*>(&ef)) {
      b.kind = 4;
      const detail::Loss& loss = lm->*exposed(detail::LimitLossTag{});
      b.alpha = detail::lossAlpha(loss);
      b.c = detail::lossC(loss);
      b.limits = &(lm->*exposed(detail::LimitLimitsTag{}));
      b.limitCount = b.limits->size();
    } else if (const auto* mp = dynamic_cast<const momentum::ModelParametersErrorFunctionT<float>*>(&ef)) {
      b.kind = 6;
      const auto& w = mp->getTargetWeights();
      const auto& t = mp->getTargetParameters().v;
      for (size_t i = 0; i < this->numParameters_; ++i) {
        b.paramWeights.push_back(i < size_t(w.size()) ? w(Eigen::Index(i)) : 0.f);
        b.targets.push_back(i < size_t(t.size()) ? t(Eigen::Index(i)) : 0.f);
      }
    } else {
      throw std::runtime_error("CudaSkeletonSolverFunction: unsupported error function class (supported: Position, Orientation, OrientationRotDiff, Plane, State, Limit, ModelParameters)");
    }
    return b;
  }
  template <class List>
  static void quats(const List& constraints, Block& b) {
    for (const auto& cst : constraints) {
      b.offsets.insert(b.offsets.end(), {cst.offset.x(), cst.offset.y(), cst.offset.z(), cst.offset.w()});
      b.targets.insert(b.targets.end(), {cst.target.x(), cst.target.y(), cst.target.z(), cst.target.w()});
    }
  }
  void addBlock(const Block& b) {
    switch (b.kind) {
      case 0: fn_->addPositionErrorFunction(b.weight, b.parents, b.offsets, b.cweights, b.alpha, b.c); break;
      case 1: fn_->addOrientationErrorFunction(b.weight, b.parents, b.offsets, b.cweights, false, b.alpha, b.c); break;
      case 2: fn_->addOrientationErrorFunction(b.weight, b.parents, b.offsets, b.cweights, true, b.alpha, b.c); break;
      case 3: fn_->addStateErrorFunction(b.weight, b.rotationErrorType ? MB2_QUATERNION_LOG_MAP : MB2_ROTATION_MATRIX_DIFFERENCE, b.posWgt, b.rotWgt, b.jointPosW, b.jointRotW); break;
      case 4: fn_->addLimitErrorFunction(b.weight, b.alpha, b.c); break;
      case 5: fn_->addPlaneErrorFunction(b.weight, b.parents, b.offsets, b.cweights, b.above, b.alpha, b.c); break;
      case 6: fn_->addModelParametersErrorFunction(b.weight, b.paramWeights); break;
      default: throw std::runtime_error("unknown block kind");
    }
  }

  const momentum::Character& character_;
  const momentum::ParameterTransformT<float>& parameterTransform_;
  int device_;
  momentum::ParameterSet enabled_;
  std::vector<ErrorFunctionPtr> errorFunctions_;
  std::vector<Block> blocks_;
  std::unique_ptr<Character> deviceCharacter_;
  std::unique_ptr<BatchedSkeletonSolverFunction> fn_;
};

// momentum::SolverT<float> over the device Gauss-Newton iteration. SolverT::solve (solver.cpp:50-128) keeps its loop, history and
// stopping rule; each doIteration() is one damped Gauss-Newton step on the GPU. solveOnDevice() hands the whole loop to the device.
class CudaGaussNewtonSolver : public momentum::SolverT<float> {
 public:
  CudaGaussNewtonSolver(const momentum::SolverOptions& options, CudaSkeletonSolverFunction* function) : momentum::SolverT<float>(options, function), fn_(function) {
    CudaGaussNewtonSolver::setOptions(options);
  }
  [[nodiscard]] std::string_view getName() const override { return "CudaGaussNewton"; }
  void setOptions(const momentum::SolverOptions& options) final { // gauss_newton_solver.cpp:37-46
    momentum::SolverT<float>::setOptions(options);
    opt_.minIterations = options.minIterations;
    opt_.maxIterations = options.maxIterations;
    opt_.threshold = options.threshold;
    opt_.verbose = options.verbose;
    if (const auto* base = dynamic_cast<const momentum::GaussNewtonSolverBaseOptions*>(&options)) {
      opt_.regularization = base->regularization;
      opt_.doLineSearch = base->doLineSearch;
    }
    if (const auto* gn = dynamic_cast<const momentum::GaussNewtonSolverOptions*>(&options)) {
      opt_.useBlockJtJ = gn->useBlockJtJ;
      opt_.targetRowsPerChunk = gn->targetRowsPerChunk;
    }
    step_.reset();
  }
  void setEnabledParameters(const momentum::ParameterSet& parameters) override {
    momentum::SolverT<float>::setEnabledParameters(parameters);
    step_.reset();
  }
  // Which of momentum's solver classes the device iteration reproduces: MB2_LINEAR_SOLVER_CHOLESKY = GaussNewtonSolverT / SubsetGaussNewtonSolverT
  // (default), MB2_LINEAR_SOLVER_QR = GaussNewtonSolverQRT (gauss_newton_solver_qr.h), MB2_LINEAR_SOLVER_TRUST_REGION_QR = TrustRegionQRT with
  // TrustRegionQROptions::trustRegionRadius_ (trust_region_qr.h:23)
  void setLinearSolver(mb2_linear_solver solver, float trustRegionRadius = 1.0f) {
    opt_.linearSolver = solver;
    opt_.trustRegionRadius = trustRegionRadius;
    step_.reset();
  }
  // the whole SolverT loop in one C-ABI call (iterations, stopping rule and error history on the device)
  double solveOnDevice(Eigen::VectorX<float>& params) {
    if (size_t(params.size()) != this->numParameters_) throw std::runtime_error("params.size() == numParameters_"); // solver.cpp:77
    GaussNewtonSolverOptions o = opt_;
    o.storeErrorHistory = true;
    BatchedGaussNewtonSolver solver(o, &fn_->batched());
    std::vector<float> p(params.data(), params.data() + params.size());
    const BatchedSolveResult r = solver.solve(p);
    for (Eigen::Index i = 0; i < params.size(); ++i) params(i) = p[size_t(i)];
    this->errorHistory_ = solver.getErrorHistory(size_t(r.iterations[0]));
    this->iteration_ = size_t(r.iterations[0]);
    this->error_ = r.errors[0];
    return r.errors[0];
  }

 protected:
  void initializeSolver() final {
    GaussNewtonSolverOptions one = opt_;
    one.minIterations = 1;
    one.maxIterations = 1;
    step_ = std::make_unique<BatchedGaussNewtonSolver>(one, &fn_->batched());
  }
  void doIteration() final { // gauss_newton_solver.cpp:224-280 on the device
    if (!step_) initializeSolver();
    std::vector<float> p(this->parameters_.data(), this->parameters_.data() + this->parameters_.size());
    const BatchedSolveResult r = step_->solve(p);
    for (Eigen::Index i = 0; i < this->parameters_.size(); ++i) this->parameters_(i) = p[size_t(i)];
    this->error_ = r.errors[0];
  }

 private:
  CudaSkeletonSolverFunction* fn_;
  GaussNewtonSolverOptions opt_;
  std::unique_ptr<BatchedGaussNewtonSolver> step_;
};

} // namespace momentum_b200
#endif
#endif

// TEST SCAFFOLDING ONLY — momentum/character_solver/state_error_function.h:17-117.
#pragma once
#include <momentum/character_solver/skeleton_error_function.h>
#include <momentum/math/transform.h>
namespace momentum {
enum class RotationErrorType { RotationMatrixDifference, QuaternionLogMap };
template <typename T>
class StateErrorFunctionT : public SkeletonErrorFunctionT<T> {
 public:
  StateErrorFunctionT(const Skeleton& skel, const ParameterTransform& pt, RotationErrorType rotationErrorType = RotationErrorType::RotationMatrixDifference)
      : SkeletonErrorFunctionT<T>(skel, pt), rotationErrorType_(rotationErrorType) {
    targetPositionWeights_ = VectorX<T>::Ones(Eigen::Index(skel.joints.size()));
    targetRotationWeights_ = VectorX<T>::Ones(Eigen::Index(skel.joints.size()));
  }
  explicit StateErrorFunctionT(const Character& character, RotationErrorType rotationErrorType = RotationErrorType::RotationMatrixDifference)
      : StateErrorFunctionT(character.skeleton, character.parameterTransform, rotationErrorType) {}
  void setTargetState(TransformListT<T> target) { targetState_ = std::move(target); }
  void setTargetWeights(const Eigen::VectorX<T>& posWeight, const Eigen::VectorX<T>& rotWeight) { targetPositionWeights_ = posWeight; targetRotationWeights_ = rotWeight; }
  void setWeights(const float posWeight, const float rotationWeight) { posWgt_ = posWeight; rotWgt_ = rotationWeight; }
  [[nodiscard]] const TransformListT<T>& getTargetState() const { return this->targetState_; }
  [[nodiscard]] const Eigen::VectorX<T>& getPositionWeights() const { return targetPositionWeights_; }
  [[nodiscard]] const Eigen::VectorX<T>& getRotationWeights() const { return targetRotationWeights_; }
  [[nodiscard]] const T& getPositionWeight() const { return posWgt_; }
  [[nodiscard]] const T& getRotationWeight() const { return rotWgt_; }

 private:
  TransformListT<T> targetState_;
  Eigen::VectorX<T> targetPositionWeights_;
  Eigen::VectorX<T> targetRotationWeights_;
  T posWgt_{1};
  T rotWgt_{1};
  const RotationErrorType rotationErrorType_;

 public:
  static constexpr T kPositionWeight = 1e-3;
  static constexpr T kOrientationWeight = 1e+0;
};
using StateErrorFunction = StateErrorFunctionT<float>;
} // namespace momentum

// TEST SCAFFOLDING ONLY — momentum/character_solver/model_parameters_error_function.h:19-66.
#pragma once
#include <momentum/character_solver/skeleton_error_function.h>
namespace momentum {
template <typename T>
class ModelParametersErrorFunctionT : public SkeletonErrorFunctionT<T> {
 public:
  ModelParametersErrorFunctionT(const Skeleton& skel, const ParameterTransform& pt) : SkeletonErrorFunctionT<T>(skel, pt) {}
  explicit ModelParametersErrorFunctionT(const Character& character) : ModelParametersErrorFunctionT(character.skeleton, character.parameterTransform) {}
  void setTargetParameters(const ModelParametersT<T>& params, const Eigen::VectorX<T>& weights) { targetParameters_ = params; targetWeights_ = weights; }
  [[nodiscard]] const ModelParametersT<T>& getTargetParameters() const { return this->targetParameters_; }
  [[nodiscard]] const Eigen::VectorX<T>& getTargetWeights() const { return this->targetWeights_; }
  static constexpr T kMotionWeight = 1e-1;

 private:
  ModelParametersT<T> targetParameters_;
  Eigen::VectorX<T> targetWeights_;
};
using ModelParametersErrorFunction = ModelParametersErrorFunctionT<float>;
} // namespace momentum

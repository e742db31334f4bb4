// TEST SCAFFOLDING ONLY — momentum/character_solver/limit_error_function.h:25-119 (limits_ and loss_ are protected there too).
#pragma once
#include <momentum/character_solver/skeleton_error_function.h>
#include <momentum/math/generalized_loss.h>
namespace momentum {
template <typename T>
class LimitErrorFunctionT : public SkeletonErrorFunctionT<T> {
 public:
  LimitErrorFunctionT(const Skeleton& skel, const ParameterTransform& pt, const ParameterLimits& pl, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : SkeletonErrorFunctionT<T>(skel, pt), limits_(pl), loss_(lossAlpha, lossC) {}
  explicit LimitErrorFunctionT(const Character& character, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : LimitErrorFunctionT(character.skeleton, character.parameterTransform, character.parameterLimits, lossAlpha, lossC) {}
  void setLimits(const ParameterLimits& lm) { limits_ = lm; }
  static constexpr T kLimitWeight = 1e1;

 protected:
  ParameterLimits limits_;
  const GeneralizedLossT<T> loss_;
};
using LimitErrorFunction = LimitErrorFunctionT<float>;
} // namespace momentum

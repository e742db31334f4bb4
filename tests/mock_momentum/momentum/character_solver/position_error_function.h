// TEST SCAFFOLDING ONLY — momentum/character_solver/position_error_function.h:16-73.
#pragma once
#include <momentum/character_solver/joint_error_function.h>
namespace momentum {
template <typename T>
struct PositionDataT : ConstraintData {
  Vector3<T> offset;
  Vector3<T> target;
  explicit PositionDataT(const Vector3<T>& inOffset, const Vector3<T>& inTarget, size_t pIndex, float w, const std::string& n = "")
      : ConstraintData(pIndex, w, n), offset(inOffset), target(inTarget) {}
};
template <typename T>
class PositionErrorFunctionT : public JointErrorFunctionT<T, PositionDataT<T>> {
 public:
  explicit PositionErrorFunctionT(const Skeleton& skel, const ParameterTransform& pt, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : JointErrorFunctionT<T, PositionDataT<T>>(skel, pt, lossAlpha, lossC) {}
  explicit PositionErrorFunctionT(const Character& character, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : PositionErrorFunctionT(character.skeleton, character.parameterTransform, lossAlpha, lossC) {}
  static constexpr T kLegacyWeight = 1e-4;
};
using PositionErrorFunction = PositionErrorFunctionT<float>;
using PositionData = PositionDataT<float>;
} // namespace momentum

// TEST SCAFFOLDING ONLY — momentum/character_solver/orientation_error_function.h:16-108.
#pragma once
#include <momentum/character_solver/joint_error_function.h>
namespace momentum {
template <typename T>
struct OrientationDataT : ConstraintData {
  Eigen::Quaternion<T> offset;
  Eigen::Quaternion<T> target;
  explicit OrientationDataT(const Eigen::Quaternion<T>& inOffset, const Eigen::Quaternion<T>& inTarget, size_t pIndex, float w, const std::string& n = "")
      : ConstraintData(pIndex, w, n), offset(inOffset.normalized()), target(inTarget.normalized()) {}
};
template <typename T>
class OrientationErrorFunctionT : public JointErrorFunctionT<T, OrientationDataT<T>, 9, 3, 0> {
 public:
  explicit OrientationErrorFunctionT(const Skeleton& skel, const ParameterTransform& pt, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : JointErrorFunctionT<T, OrientationDataT<T>, 9, 3, 0>(skel, pt, lossAlpha, lossC) {}
  explicit OrientationErrorFunctionT(const Character& character, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : OrientationErrorFunctionT(character.skeleton, character.parameterTransform, lossAlpha, lossC) {}
  static constexpr T kLegacyWeight = 1e-1;
};
template <typename T>
class OrientationRotDiffErrorFunctionT : public JointErrorFunctionT<T, OrientationDataT<T>, 9, 3, 0> {
 public:
  explicit OrientationRotDiffErrorFunctionT(const Skeleton& skel, const ParameterTransform& pt, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : JointErrorFunctionT<T, OrientationDataT<T>, 9, 3, 0>(skel, pt, lossAlpha, lossC) {}
  explicit OrientationRotDiffErrorFunctionT(const Character& character, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : OrientationRotDiffErrorFunctionT(character.skeleton, character.parameterTransform, lossAlpha, lossC) {}
  static constexpr T kLegacyWeight = 1e-1;
};
using OrientationErrorFunction = OrientationErrorFunctionT<float>;
using OrientationData = OrientationDataT<float>;
} // namespace momentum

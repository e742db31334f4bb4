// TEST SCAFFOLDING ONLY — momentum/character_solver/plane_error_function.h:18-101 (halfPlane_ is private there too).
#pragma once
#include <momentum/character_solver/joint_error_function.h>
namespace momentum {
template <typename T>
struct PlaneDataT : ConstraintData {
  Vector3<T> offset;
  Vector3<T> normal;
  T d;
  explicit PlaneDataT(const Vector3<T>& inOffset, const Vector3<T>& inNormal, const T inD, size_t pIndex, float w, const std::string& n = "")
      : ConstraintData(pIndex, w, n), offset(inOffset), normal(inNormal.normalized()), d(inD) {}
};
template <typename T>
class PlaneErrorFunctionT : public JointErrorFunctionT<T, PlaneDataT<T>, 1> {
 public:
  explicit PlaneErrorFunctionT(const Skeleton& skel, const ParameterTransform& pt, const bool above = false, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : JointErrorFunctionT<T, PlaneDataT<T>, 1>(skel, pt, lossAlpha, lossC), halfPlane_(above) {}
  explicit PlaneErrorFunctionT(const Character& character, const bool above = false, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : PlaneErrorFunctionT(character.skeleton, character.parameterTransform, above, lossAlpha, lossC) {}
  static constexpr T kLegacyWeight = 1e-4;

 private:
  bool halfPlane_;
};
using PlaneErrorFunction = PlaneErrorFunctionT<float>;
using PlaneData = PlaneDataT<float>;
} // namespace momentum

// TEST SCAFFOLDING ONLY — momentum/math/transform.h:36-42 (data members only).
#pragma once
#include <momentum/math/types.h>
namespace momentum {
template <class T>
struct TransformT {
  Quaternion<T> rotation = Quaternion<T>::Identity();
  Vector3<T> translation;
  T scale = T(1);
};
template <class T> using TransformListT = std::vector<TransformT<T>>;
} // namespace momentum

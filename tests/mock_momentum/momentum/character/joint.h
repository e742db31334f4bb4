// TEST SCAFFOLDING ONLY — momentum/character/joint.h:18-76 (fields).
#pragma once
#include <momentum/character/types.h>
namespace momentum {
template <class T>
struct JointT {
  std::string name;
  size_t parent = kInvalidIndex;
  Quaternion<T> preRotation = Quaternion<T>::Identity();
  Vector3<T> translationOffset;
};
using Joint = JointT<float>;
using JointList = std::vector<Joint>;
} // namespace momentum

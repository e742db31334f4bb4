// TEST SCAFFOLDING ONLY — momentum/character/skeleton.h:22-77 (the joint list).
#pragma once
#include <momentum/character/joint.h>
namespace momentum {
template <class T>
struct SkeletonT {
  JointList joints; // stored as float whatever T is (skeleton.h:25)
};
using Skeleton = SkeletonT<float>;
} // namespace momentum

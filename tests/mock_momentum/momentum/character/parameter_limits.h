// TEST SCAFFOLDING ONLY — momentum/character/parameter_limits.h:20-138 (same member names; the union is a plain struct here).
#pragma once
#include <momentum/character/types.h>
namespace momentum {
enum LimitType { MinMax, MinMaxJoint, MinMaxJointPassive, Linear, LinearJoint, Ellipsoid, HalfPlane };
struct LimitMinMax { size_t parameterIndex; Vector2f limits; };
struct LimitMinMaxJoint { size_t jointIndex; size_t jointParameter; Vector2f limits; };
struct LimitLinear { size_t referenceIndex; size_t targetIndex; float scale; float offset; float rangeMin; float rangeMax; };
struct LimitLinearJoint { size_t referenceJointIndex; size_t referenceJointParameter; size_t targetJointIndex; size_t targetJointParameter; float scale; float offset; float rangeMin; float rangeMax; };
struct LimitEllipsoid { Affine3f ellipsoid; Affine3f ellipsoidInv; Vector3f offset; size_t ellipsoidParent; size_t parent; };
struct LimitHalfPlane { size_t param1; size_t param2; Vector2f normal; float offset; };
struct LimitData {
  LimitMinMax minMax{};
  LimitMinMaxJoint minMaxJoint{};
  LimitLinear linear{};
  LimitLinearJoint linearJoint{};
  LimitEllipsoid ellipsoid{};
  LimitHalfPlane halfPlane{};
};
struct ParameterLimit {
  LimitData data;
  LimitType type = LimitType::MinMax;
  float weight = 1.0f;
};
using ParameterLimits = std::vector<ParameterLimit>;
} // namespace momentum

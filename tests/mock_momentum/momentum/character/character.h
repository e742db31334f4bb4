// TEST SCAFFOLDING ONLY — momentum/character/character.h:32-46 (the three members the IK path uses).
#pragma once
#include <momentum/character/parameter_limits.h>
#include <momentum/character/parameter_transform.h>
#include <momentum/character/skeleton.h>
namespace momentum {
struct Character {
  Skeleton skeleton;
  ParameterTransform parameterTransform;
  ParameterLimits parameterLimits;
};
} // namespace momentum

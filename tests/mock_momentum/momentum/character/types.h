// TEST SCAFFOLDING ONLY — momentum/character/types.h:21.
#pragma once
#include <momentum/math/types.h>
namespace momentum {
inline constexpr size_t kParametersPerJoint = 7;
}

// TEST SCAFFOLDING ONLY — momentum/character_solver/error_function_types.h:34-44.
#pragma once
#include <momentum/math/types.h>
namespace momentum {
struct ConstraintData {
  size_t parent = kInvalidIndex;
  float weight = 0.0f;
  std::string name = {};
  ConstraintData(size_t pIndex, float w, const std::string& n = "") : parent(pIndex), weight(w), name(n) {}
};
} // namespace momentum

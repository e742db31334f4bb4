// TEST SCAFFOLDING ONLY — momentum/character/parameter_transform.h:62-184 (fields the solver path reads).
#pragma once
#include <momentum/character/types.h>
namespace momentum {
template <class T>
struct ParameterTransformT {
  std::vector<std::string> name;
  SparseRowMatrix<T> transform; // (7 * joints) x parameters
  VectorX<T> offsets;
  VectorX<bool> activeJointParams;
  [[nodiscard]] Eigen::Index numAllModelParameters() const { return transform.cols(); }
  [[nodiscard]] Eigen::Index numJointParameters() const { return transform.rows(); }
};
using ParameterTransform = ParameterTransformT<float>;
} // namespace momentum

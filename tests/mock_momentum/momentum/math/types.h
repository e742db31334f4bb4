// TEST SCAFFOLDING ONLY — stand-in for momentum/math/types.h: the aliases the solver interfaces use (math/types.h:62-100,426-429).
#pragma once
#include <Eigen/Core>
#include <bitset>
#include <cstddef>
#include <limits>
#include <string>
#include <vector>
namespace momentum {
template <class T> using VectorX = Eigen::VectorX<T>;
template <class T> using MatrixX = Eigen::MatrixX<T>;
template <class T> using Vector3 = Eigen::Vector3<T>;
template <class T> using Quaternion = Eigen::Quaternion<T>;
template <class T> using SparseRowMatrix = Eigen::SparseMatrix<T, Eigen::RowMajor>;
template <class T> using SparseMatrix = Eigen::SparseMatrix<T>;
using Vector3f = Eigen::Vector3f;
using Vector2f = Eigen::Vector2f;
using VectorXf = Eigen::VectorXf;
using VectorXi = Eigen::VectorXi;
using MatrixXf = Eigen::MatrixXf;
using Quaternionf = Eigen::Quaternionf;
using Affine3f = Eigen::Affine3f;
inline constexpr size_t kInvalidIndex = std::numeric_limits<size_t>::max();
inline constexpr size_t kMaxModelParams = 2048;
using ParameterSet = std::bitset<kMaxModelParams>;
template <class T>
struct ModelParametersT { // strong typedef around the parameter vector (math/types.h EigenStrongType)
  VectorX<T> v;
  ModelParametersT() = default;
  ModelParametersT(const VectorX<T>& x) : v(x) {}
  [[nodiscard]] Eigen::Index size() const { return v.size(); }
  T& operator()(Eigen::Index i) { return v(i); }
  const T& operator()(Eigen::Index i) const { return v(i); }
};
using ModelParameters = ModelParametersT<float>;
} // namespace momentum

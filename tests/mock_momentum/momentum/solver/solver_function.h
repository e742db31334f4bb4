// TEST SCAFFOLDING ONLY — the virtual interface of momentum/solver/solver_function.h:36-192, signature for signature, so that a
// subclass written against momentum compiles here. Non-pure defaults do the least that keeps the contract.
#pragma once
#include <momentum/math/types.h>
#include <momentum/solver/fwd.h>
#include <string>
#include <unordered_map>
namespace momentum {
inline constexpr size_t kJacobianRowAlignment = 8;
[[nodiscard]] constexpr size_t padToSimdAlignment(size_t size) { return (size + kJacobianRowAlignment - 1) & ~(kJacobianRowAlignment - 1); }
template <typename T>
class SolverFunctionT {
 public:
  virtual ~SolverFunctionT() = default;
  virtual double getError(const VectorX<T>& parameters) = 0;
  virtual double getGradient(const VectorX<T>& parameters, VectorX<T>& gradient) = 0;
  virtual void getHessian(const VectorX<T>&, MatrixX<T>&) {}
  virtual double getJtJR(const VectorX<T>& parameters, MatrixX<T>& jtj, VectorX<T>& jtr) {
    (void)parameters; (void)jtj; (void)jtr;
    return 0.0; // (the reference's default accumulates it block-wise from computeJacobianBlock, solver_function.cpp:74-121)
  }
  virtual double getJtJR_Sparse(const VectorX<T>&, SparseMatrix<T>&, VectorX<T>&) { return 0.0; }
  virtual void initializeJacobianComputation(const VectorX<T>& parameters) = 0;
  [[nodiscard]] virtual size_t getJacobianBlockCount() const = 0;
  [[nodiscard]] virtual size_t getJacobianBlockSize(size_t blockIndex) const = 0;
  virtual double computeJacobianBlock(const VectorX<T>& parameters, size_t blockIndex, Eigen::Ref<MatrixX<T>> jacobianBlock,
                                      Eigen::Ref<VectorX<T>> residualBlock, size_t& actualRows) = 0;
  virtual void finalizeJacobianComputation() {}
  virtual double getSolverDerivatives(const VectorX<T>& parameters, MatrixX<T>& hess, VectorX<T>& grad) { return getJtJR(parameters, hess, grad); }
  virtual void updateParameters(VectorX<T>& parameters, const VectorX<T>& gradient) = 0;
  virtual void setEnabledParameters(const ParameterSet&) {}
  [[nodiscard]] size_t getNumParameters() const { return numParameters_; }
  [[nodiscard]] size_t getActualParameters() const { return actualParameters_; }
  virtual void storeHistory(std::unordered_map<std::string, MatrixX<T>>&, size_t, size_t) {}

 protected:
  size_t numParameters_{};
  size_t actualParameters_{};
};
} // namespace momentum

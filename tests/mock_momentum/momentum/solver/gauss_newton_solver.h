// TEST SCAFFOLDING ONLY — momentum/solver/gauss_newton_solver.h:17-137 reduced to the useBlockJtJ branch of doIteration
// (gauss_newton_solver.cpp:69-107,224-259: getJtJR, in-place compaction to the enabled parameters, (JtJ + lambda I) delta = Jtr by a
// plain dense Cholesky, updateParameters): enough to show the STOCK solver loop driving a device-backed SolverFunctionT unmodified.
#pragma once
#include <momentum/solver/solver.h>
#include <cstdint>
namespace momentum {
struct GaussNewtonSolverBaseOptions : SolverOptions {
  float regularization = 0.05f;
  bool doLineSearch = false;
  GaussNewtonSolverBaseOptions() = default;
  /* implicit */ GaussNewtonSolverBaseOptions(const SolverOptions& baseOptions) : SolverOptions(baseOptions) {}
};
struct GaussNewtonSolverOptions : GaussNewtonSolverBaseOptions {
  bool useBlockJtJ = false;
  size_t targetRowsPerChunk = SIZE_MAX;
  GaussNewtonSolverOptions() = default;
  /* implicit */ GaussNewtonSolverOptions(const SolverOptions& baseOptions) : GaussNewtonSolverBaseOptions(baseOptions) {}
};
template <typename T>
class GaussNewtonSolverT : public SolverT<T> {
 public:
  GaussNewtonSolverT(const SolverOptions& options, SolverFunctionT<T>* solver) : SolverT<T>(options, solver) { GaussNewtonSolverT::setOptions(options); }
  [[nodiscard]] std::string_view getName() const override { return "GaussNewton"; }
  void setOptions(const SolverOptions& options) final {
    SolverT<T>::setOptions(options);
    if (const auto* o = dynamic_cast<const GaussNewtonSolverOptions*>(&options)) { regularization_ = o->regularization; useBlockJtJ_ = o->useBlockJtJ; }
  }

 protected:
  void initializeSolver() final {}
  void doIteration() final {
    if (!useBlockJtJ_) throw std::runtime_error("scaffolding: only the useBlockJtJ path is restated");
    MatrixX<T> jtj;
    VectorX<T> jtr;
    this->error_ = this->solverFunction_->getJtJR(this->parameters_, jtj, jtr);
    std::vector<int> en;
    for (size_t i = 0; i < this->numParameters_; ++i) if (this->activeParameters_.test(i)) en.push_back(int(i));
    const int n = int(en.size());
    std::vector<double> A(size_t(n) * n, 0.0), b(n);
    for (int a = 0; a < n; ++a) { b[a] = jtr(en[a]); for (int c = 0; c <= a; ++c) A[size_t(a) * n + c] = jtj(en[a], en[c]); A[size_t(a) * n + a] += regularization_; }
    for (int k = 0; k < n; ++k) { // dense LLT, then the two triangular solves
      double x = A[size_t(k) * n + k];
      for (int j = 0; j < k; ++j) x -= A[size_t(k) * n + j] * A[size_t(k) * n + j];
      x = std::sqrt(x);
      A[size_t(k) * n + k] = x;
      for (int i = k + 1; i < n; ++i) { double s = A[size_t(i) * n + k]; for (int j = 0; j < k; ++j) s -= A[size_t(i) * n + j] * A[size_t(k) * n + j]; A[size_t(i) * n + k] = s / x; }
    }
    for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[size_t(i) * n + k] * b[k]; b[i] = s / A[size_t(i) * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[size_t(k) * n + i] * b[k]; b[i] = s / A[size_t(i) * n + i]; }
    VectorX<T> delta = VectorX<T>::Zero(Eigen::Index(this->numParameters_));
    for (int a = 0; a < n; ++a) delta(en[a]) = T(b[a]);
    this->solverFunction_->updateParameters(this->parameters_, delta);
  }

 private:
  float regularization_ = 0.05f;
  bool useBlockJtJ_ = false;
};
} // namespace momentum

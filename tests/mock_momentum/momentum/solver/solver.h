// TEST SCAFFOLDING ONLY — momentum/solver/solver.h:19-134 with the iteration driver of solver.cpp:50-128 (history matrices left out),
// so that a SolverT subclass is exercised through the same solve() a momentum build would run.
#pragma once
#include <momentum/solver/solver_function.h>
#include <cmath>
#include <limits>
#include <stdexcept>
#include <string_view>
namespace momentum {
struct SolverOptions {
  size_t minIterations = 1;
  size_t maxIterations = 2;
  float threshold = 1.0f;
  bool verbose = false;
  virtual ~SolverOptions() = default;
};
template <typename T>
class SolverT {
 public:
  SolverT(const SolverOptions& options, SolverFunctionT<T>* solver) : solverFunction_(solver) {
    numParameters_ = solver->getNumParameters();
    activeParameters_.flip();
    actualParameters_ = int(numParameters_);
    SolverT::setOptions(options);
  }
  virtual ~SolverT() = default;
  [[nodiscard]] virtual std::string_view getName() const = 0;
  virtual void setOptions(const SolverOptions& options) {
    minIterations_ = options.minIterations; maxIterations_ = options.maxIterations; threshold_ = options.threshold; verbose_ = options.verbose;
  }
  double solve(Eigen::VectorX<T>& params) {
    errorHistory_.clear();
    if (size_t(params.size()) != numParameters_) throw std::runtime_error("params.size() == numParameters_");
    parameters_ = params;
    error_ = lastError_ = std::numeric_limits<double>::max();
    initializeSolver();
    for (iteration_ = 0; iteration_ < maxIterations_; iteration_++) {
      doIteration();
      errorHistory_.push_back(error_);
      const bool converged = std::fabs(lastError_ - error_) / (std::fabs(error_) + std::numeric_limits<float>::min()) <= threshold_ * std::numeric_limits<float>::epsilon();
      if (iteration_ >= minIterations_ && converged) break;
      lastError_ = error_;
    }
    params = parameters_;
    return error_;
  }
  virtual void setEnabledParameters(const ParameterSet& parameters) {
    activeParameters_ = parameters;
    actualParameters_ = 0;
    for (size_t i = 0; i < numParameters_; ++i) if (parameters.test(i)) actualParameters_++;
    newParameterPattern_ = true;
    solverFunction_->setEnabledParameters(parameters);
  }
  [[nodiscard]] const ParameterSet& getActiveParameters() const { return activeParameters_; }
  void setParameters(const Eigen::VectorX<T>& params) { parameters_ = params; }
  [[nodiscard]] size_t getMinIterations() const { return minIterations_; }
  [[nodiscard]] size_t getMaxIterations() const { return maxIterations_; }
  [[nodiscard]] size_t getNumParameters() const { return numParameters_; }
  [[nodiscard]] const std::vector<double>& getErrorHistory() const { return errorHistory_; }

 protected:
  virtual void initializeSolver() = 0;
  virtual void doIteration() = 0;
  size_t numParameters_;
  SolverFunctionT<T>* solverFunction_;
  Eigen::VectorX<T> parameters_;
  ParameterSet activeParameters_;
  int actualParameters_;
  bool newParameterPattern_ = true;
  size_t iteration_{};
  double error_{};
  double lastError_{};
  std::vector<double> errorHistory_;
  size_t minIterations_{}, maxIterations_{};
  float threshold_{};
  bool verbose_{};
};
} // namespace momentum

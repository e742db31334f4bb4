// TEST SCAFFOLDING ONLY — momentum/character_solver/joint_error_function.h:54-221: constraint list + loss, same member names and access.
#pragma once
#include <momentum/character_solver/error_function_types.h>
#include <momentum/character_solver/skeleton_error_function.h>
#include <momentum/math/generalized_loss.h>
#include <span>
namespace momentum {
template <typename T, class Data, size_t FuncDim = 3, size_t NumVec = 1, size_t NumPos = 1>
class JointErrorFunctionT : public SkeletonErrorFunctionT<T> {
 public:
  static constexpr size_t kFuncDim = FuncDim;
  JointErrorFunctionT(const Skeleton& skel, const ParameterTransform& pt, const T& lossAlpha = GeneralizedLossT<T>::kL2, const T& lossC = T(1))
      : SkeletonErrorFunctionT<T>(skel, pt), loss_(lossAlpha, lossC) {}
  [[nodiscard]] size_t getJacobianSize() const final { return FuncDim * constraints_.size(); }
  void addConstraint(const Data& constr) { constraints_.push_back(constr); }
  void addConstraints(std::span<const Data> constrs) { constraints_.insert(constraints_.end(), constrs.begin(), constrs.end()); }
  void setConstraints(std::span<const Data> constrs) { constraints_.assign(constrs.begin(), constrs.end()); }
  [[nodiscard]] const std::vector<Data>& getConstraints() const { return constraints_; }
  [[nodiscard]] size_t getNumConstraints() const { return constraints_.size(); }
  void clearConstraints() { constraints_.clear(); }
  Data& getConstraint(size_t index) { return constraints_.at(index); }

 protected:
  std::vector<Data> constraints_;
  const GeneralizedLossT<T> loss_;
};
} // namespace momentum

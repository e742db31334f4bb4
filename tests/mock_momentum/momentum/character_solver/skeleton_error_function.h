// TEST SCAFFOLDING ONLY — momentum/character_solver/skeleton_error_function.h:19-150: the state an adapter translates. The
// evaluation virtuals (getError / getGradient / getJacobian) are omitted: the device path replaces them.
#pragma once
#include <momentum/character/character.h>
#include <momentum/character/skeleton_state.h>
namespace momentum {
template <typename T>
class SkeletonErrorFunctionT {
 public:
  SkeletonErrorFunctionT(const Skeleton& skel, const ParameterTransform& pt) : skeleton_(skel), parameterTransform_(pt), weight_(1.0), activeJointParams_(pt.activeJointParams) {
    enabledParameters_.flip();
  }
  virtual ~SkeletonErrorFunctionT() = default;
  [[nodiscard]] const Skeleton& getSkeleton() const { return skeleton_; }
  [[nodiscard]] const ParameterTransform& getParameterTransform() const { return parameterTransform_; }
  void setWeight(T w) { weight_ = w; }
  [[nodiscard]] T getWeight() const { return weight_; }
  void setActiveJoints(const VectorX<bool>& aj) { activeJointParams_ = aj; }
  void setEnabledParameters(const ParameterSet& ps) { enabledParameters_ = ps; }
  [[nodiscard]] virtual size_t getJacobianSize() const { return 0; }

 protected:
  const Skeleton& skeleton_;
  const ParameterTransform& parameterTransform_;
  T weight_;
  VectorX<bool> activeJointParams_;
  ParameterSet enabledParameters_;
};
using SkeletonErrorFunction = SkeletonErrorFunctionT<float>;
} // namespace momentum

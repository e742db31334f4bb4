// TEST SCAFFOLDING ONLY — momentum/math/generalized_loss.h:46-110 (the constants and state an adapter reads).
#pragma once
#include <cmath>
#include <limits>
namespace momentum {
template <class T>
class GeneralizedLossT {
 public:
  static constexpr T kL2 = T(2);
  static constexpr T kL1 = T(1);
  static constexpr T kCauchy = T(0);
  static constexpr T kWelsch = std::numeric_limits<T>::lowest();
  GeneralizedLossT(const T& a = kL2, const T& c = T(1)) : alpha_(a), invC2_(T(1) / (c * c)) {}
  [[nodiscard]] bool isL2() const { return alpha_ == kL2; }
  [[nodiscard]] T invC2() const { return invC2_; }

 protected:
  const T alpha_;
  const T invC2_;
};
} // namespace momentum

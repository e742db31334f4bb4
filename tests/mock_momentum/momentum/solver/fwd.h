// TEST SCAFFOLDING ONLY — momentum/solver/fwd.h.
#pragma once
namespace momentum {
template <class T> class SolverFunctionT;
template <class T> class SolverT;
} // namespace momentum

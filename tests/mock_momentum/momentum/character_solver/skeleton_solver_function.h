// TEST SCAFFOLDING ONLY — presence marker: the adapters header looks for this file to enable its momentum-facing part.
#pragma once
#include <momentum/character_solver/skeleton_error_function.h>
#include <momentum/solver/solver_function.h>

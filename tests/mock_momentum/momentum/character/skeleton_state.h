// TEST SCAFFOLDING ONLY — forward declarations (skeleton_state.h / mesh_state.h are not needed by the adapters).
#pragma once
namespace momentum {
template <class T> struct SkeletonStateT;
template <class T> struct MeshStateT;
} // namespace momentum

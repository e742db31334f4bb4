// Per-frame marker tracking on the batched device solver: the C++ caller that turns a mocap sequence into a batch of IK instances
// (SURVEY 8(f) rank 2). Mirrors momentum::trackPosesForFrames (marker_tracking/marker_tracker.cpp:848-1051, rigid warm start
// :801-831) over the C++ host side of momentum_b200_adapters.hpp (part 1: no momentum / Eigen types needed):
//
//   objective per frame (marker_tracker.cpp:917-969)   Limit (weight 0.1) + Position over the visible markers (kLegacyWeight * markerWeight,
//                                                      lossAlpha) + half-plane floor constraints on the "Floor_" locators (weight 5 * locator
//                                                      weight, y-up, y = 0, kLegacyWeight) [+ ModelParameters smoothness to the previous pose
//                                                      when continuous]
//   rigid warm start (:801-831)                        only the rigid parameters enabled, up to 50 iterations, smoothing off
//   full solve (:1030-1036)                            the pose parameters enabled, minIterations 2, maxIterations config.maxIter, threshold 1
//
// isContinuous = false (every frame starts from its own column of initialMotion: the calibration keyframe mode) is embarrassingly parallel:
// all valid frames become ONE batch — a frame's visible markers are per-instance constraint weights over the full locator list (an
// occluded marker has weight 0, which the reference expresses by not adding the constraint; joint_error_function-inl.h:197-199 skips
// zero weights) — solved in two batched calls (rigid, then pose). isContinuous = true carries the previous frame's solution and is
// sequential by construction: one instance per call (the persistent single-launch kernel serves those).
// The reference solves with GaussNewtonSolverQRT; this front end uses the device Gauss-Newton (same normal equations, Cholesky).
// Out of scope here as on the device path: skinned locators, collision, gloves, gap filling (pass pre-processed markers).
#pragma once

#include <algorithm>
#include <map>
#include <optional>
#include <string>
#include <vector>

#include "momentum_b200_adapters.hpp"

namespace momentum_b200 {

struct Marker { // momentum/character/marker.h:18-31
  std::string name = "Undefined";
  double pos[3] = {0.0, 0.0, 0.0};
  bool occluded = true;
  float confidence = 1.0f;
};

struct Locator { // the fields of momentum::Locator the tracker reads (name, parent joint, offset in the parent frame, weight)
  std::string name;
  int32_t parent = 0;
  float offset[3] = {0.f, 0.f, 0.f};
  float weight = 1.f;
};

struct TrackingConfig { // marker_tracker.h:42-55,95-120 (the fields this path uses)
  float minVisPercent = 0.f;
  float lossAlpha = 2.0f;
  size_t maxIter = 30;
  float regularization = 0.05f;
  bool debug = false;
  float smoothing = 0.f;
  float markerWeight = 1.0f;
  std::optional<ParameterSet> activeParams;
};

struct TrackingResult {
  std::vector<float> motion;      // [frames][parameters] (= MatrixXf parameters x frames, column-major)
  size_t solvedFrames = 0;
  double priorError = 0.0, error = 0.0; // sums over the solved frames (marker_tracker.cpp:1034-1035)
};

namespace detail {
constexpr float kPositionLegacyWeight = 1e-4f; // position_error_function.h:64
constexpr float kPlaneLegacyWeight = 1e-4f;    // plane_error_function.h:83
constexpr float kLimitWeight = 0.1f;           // marker_tracker.cpp:916
constexpr float kFloorWeight = 5.0f;           // :933
} // namespace detail

// markerData [frames][markers]; initialMotion [frames][numParameters]; rigidParameters = ParameterTransform::getRigidParameters(),
// poseParameters = getPoseParameters() minus the "locators" set (marker_tracker.cpp:896-911 does that intersection at the call site).
inline TrackingResult trackPosesForFrames(const std::vector<std::vector<Marker>>& markerData, const Character& character, size_t numParameters,
                                          const std::vector<Locator>& locators, const std::vector<float>& initialMotion, const TrackingConfig& config,
                                          const std::vector<size_t>& frameIndices, bool isContinuous, const ParameterSet& rigidParameters,
                                          const ParameterSet& poseParametersIn) {
  const size_t numFrames = markerData.size(), n = numParameters;
  if (numFrames == 0) throw std::runtime_error("Input marker data is empty.");
  if (initialMotion.size() < numFrames * n) throw std::runtime_error("Number of frames in data exceeds input motion columns");
  std::vector<size_t> sortedFrames = frameIndices;
  std::sort(sortedFrames.begin(), sortedFrames.end());
  ParameterSet poseParams = poseParametersIn;
  if (config.activeParams) poseParams &= *config.activeParams;

  // constraint topology shared by every frame: all locators (markers are matched by name, createConstraintData tracker_utils.cpp:37-70)
  std::map<std::string, size_t> locatorLookup;
  for (size_t i = 0; i < locators.size(); ++i) locatorLookup[locators[i].name] = i;
  const size_t nl = locators.size();
  std::vector<int32_t> posParents(nl);
  std::vector<float> posOffsets(3 * nl), ones(std::max<size_t>(nl, 1), 1.f);
  for (size_t i = 0; i < nl; ++i) { posParents[i] = locators[i].parent; for (int k = 0; k < 3; ++k) posOffsets[3 * i + k] = locators[i].offset[k]; }
  std::vector<int32_t> floorParents;
  std::vector<float> floorOffsets, floorWeights;
  for (const auto& loc : locators)
    if (loc.name.rfind("Floor_", 0) == 0) { // createFloorConstraints (plane_error_function.cpp:15-35)
      floorParents.push_back(loc.parent);
      floorOffsets.insert(floorOffsets.end(), loc.offset, loc.offset + 3);
      floorWeights.push_back(loc.weight * detail::kFloorWeight);
    }
  // per frame: weights (0 = marker not visible / not mapped) and targets over the locator list
  auto frameConstraints = [&](size_t iFrame, std::vector<float>& w, std::vector<float>& t) -> size_t {
    w.assign(nl, 0.f);
    t.assign(3 * nl, 0.f);
    size_t visible = 0;
    for (const auto& m : markerData[iFrame]) {
      if (m.occluded) continue;
      const auto q = locatorLookup.find(m.name);
      if (q == locatorLookup.end()) continue;
      const size_t li = q->second;
      w[li] = locators[li].weight * m.confidence;
      for (int k = 0; k < 3; ++k) t[3 * li + k] = float(m.pos[k]);
      ++visible;
    }
    return visible;
  };
  auto isValid = [&](size_t iFrame, size_t visible) { return float(visible) > float(markerData[iFrame].size()) * config.minVisPercent; }; // :1005

  auto makeFunction = [&](int32_t batch, int& posIdx, int& planeIdx, int& smoothIdx, bool smooth) {
    auto fn = std::make_unique<BatchedSkeletonSolverFunction>(character, batch);
    fn->addLimitErrorFunction(detail::kLimitWeight);
    posIdx = fn->addPositionErrorFunction(detail::kPositionLegacyWeight * config.markerWeight, posParents, posOffsets, ones, config.lossAlpha, 1.f);
    planeIdx = floorParents.empty() ? -1 : fn->addPlaneErrorFunction(detail::kPlaneLegacyWeight, floorParents, floorOffsets, floorWeights, /*above=*/true);
    smoothIdx = -1;
    if (smooth) { // ModelParametersErrorFunction(character, poseParams & ~rigid) (:957-961)
      std::vector<float> tw(n, 0.f);
      for (size_t i = 0; i < n; ++i) tw[i] = (poseParams.test(i) && !rigidParameters.test(i)) ? 1.f : 0.f;
      smoothIdx = fn->addModelParametersErrorFunction(config.smoothing, tw);
    }
    return fn;
  };
  auto floorTargets = [&](int32_t batch) {
    std::vector<float> planes(size_t(batch) * floorParents.size() * 4, 0.f);
    for (size_t i = 0; i < planes.size() / 4; ++i) planes[4 * i + 1] = 1.f; // normal = UnitY, d = 0
    return planes;
  };
  GaussNewtonSolverOptions rigidOpt, fullOpt; // :913-920 and solveRigidInitialization :813-814
  fullOpt.maxIterations = config.maxIter;
  fullOpt.minIterations = 2;
  fullOpt.doLineSearch = false;
  fullOpt.threshold = 1.f;
  fullOpt.regularization = config.regularization;
  rigidOpt = fullOpt;
  rigidOpt.maxIterations = 50;

  TrackingResult result;
  result.motion = initialMotion;
  std::vector<float> w, t;

  if (!isContinuous) {
    // ---- every valid frame is an instance of one batch ----
    std::vector<size_t> valid;
    std::vector<float> W, T;
    for (const size_t iFrame : sortedFrames) {
      if (iFrame >= numFrames) throw std::runtime_error("frame index out of range");
      const size_t visible = frameConstraints(iFrame, w, t);
      if (!isValid(iFrame, visible)) continue;
      valid.push_back(iFrame);
      W.insert(W.end(), w.begin(), w.end());
      T.insert(T.end(), t.begin(), t.end());
    }
    std::vector<float> dof(valid.size() * n);
    for (size_t s = 0; s < valid.size(); ++s) std::copy(initialMotion.begin() + valid[s] * n, initialMotion.begin() + (valid[s] + 1) * n, dof.begin() + s * n);
    if (!valid.empty()) {
      const int32_t B = int32_t(valid.size());
      int posIdx, planeIdx, smoothIdx;
      auto fn = makeFunction(B, posIdx, planeIdx, smoothIdx, false);
      if (nl > 0) { fn->setConstraintWeights(posIdx, W, true); fn->setTargets(posIdx, T); }
      if (planeIdx >= 0) fn->setTargets(planeIdx, floorTargets(B));
      // rigid warm start for every frame (needsInit is set per frame in this mode, :1000-1003)
      fn->setEnabledParameters(rigidParameters);
      { BatchedGaussNewtonSolver solver(rigidOpt, fn.get()); solver.solve(dof); }
      fn->setEnabledParameters(poseParams);
      const std::vector<double> prior = fn->getError(dof);
      BatchedGaussNewtonSolver solver(fullOpt, fn.get());
      const BatchedSolveResult r = solver.solve(dof);
      for (int32_t s = 0; s < B; ++s) { result.priorError += prior[size_t(s)]; result.error += r.errors[size_t(s)]; }
      result.solvedFrames = valid.size();
    }
    // store (:1039-1048): every frame up to a solved / visited frame takes that frame's pose; the tail takes the last one
    size_t outputIndex = 0, nextValid = 0;
    const float* last = initialMotion.data() + (sortedFrames.empty() ? 0 : sortedFrames[0]) * n;
    std::vector<float> hold(n);
    for (const size_t iFrame : sortedFrames) {
      if (nextValid < valid.size() && valid[nextValid] == iFrame) { last = dof.data() + nextValid * n; ++nextValid; }
      else last = initialMotion.data() + iFrame * n; // an unsolved frame keeps its initial column (dof = initialMotion.col(iFrame), :1001)
      std::copy(last, last + n, hold.begin());
      while (outputIndex <= iFrame) { std::copy(hold.begin(), hold.end(), result.motion.begin() + outputIndex * n); ++outputIndex; }
    }
    std::copy(last, last + n, hold.begin());
    while (outputIndex < numFrames) { std::copy(hold.begin(), hold.end(), result.motion.begin() + outputIndex * n); ++outputIndex; }
    return result;
  }

  // ---- continuous: the previous frame's solution is the start (and the smoothness target) of the next: one instance per call ----
  int posIdx, planeIdx, smoothIdx;
  auto fn = makeFunction(1, posIdx, planeIdx, smoothIdx, true);
  if (planeIdx >= 0) fn->setTargets(planeIdx, floorTargets(1));
  std::vector<float> dof(initialMotion.begin() + (sortedFrames.empty() ? 0 : sortedFrames[0]) * n, initialMotion.begin() + ((sortedFrames.empty() ? 0 : sortedFrames[0]) + 1) * n);
  bool needsInit = true; // isGlobalTransformZero (:732-739, :990)
  for (size_t i = 0; i < n; ++i) if (rigidParameters.test(i) && dof[i] != 0.f) needsInit = false;
  size_t outputIndex = 0;
  for (const size_t iFrame : sortedFrames) {
    if (iFrame >= numFrames) throw std::runtime_error("frame index out of range");
    const size_t visible = frameConstraints(iFrame, w, t);
    if (isValid(iFrame, visible)) {
      if (nl > 0) { fn->setConstraintWeights(posIdx, w, true); fn->setTargets(posIdx, t); }
      if (needsInit) { // solveRigidInitialization
        fn->setErrorFunctionWeight(smoothIdx, 0.f);
        fn->setEnabledParameters(rigidParameters);
        fn->setTargets(smoothIdx, dof);
        BatchedGaussNewtonSolver solver(rigidOpt, fn.get());
        solver.solve(dof);
        fn->setErrorFunctionWeight(smoothIdx, config.smoothing);
        needsInit = false;
      }
      fn->setEnabledParameters(poseParams);
      fn->setTargets(smoothIdx, dof); // smoothness target = the last pose (:1025-1027)
      result.priorError += fn->getError(dof)[0];
      BatchedGaussNewtonSolver solver(fullOpt, fn.get());
      result.error += solver.solve(dof).errors[0];
      ++result.solvedFrames;
    }
    while (outputIndex <= iFrame) { std::copy(dof.begin(), dof.end(), result.motion.begin() + outputIndex * n); ++outputIndex; }
  }
  while (outputIndex < numFrames) { std::copy(dof.begin(), dof.end(), result.motion.begin() + outputIndex * n); ++outputIndex; }
  return result;
}

} // namespace momentum_b200

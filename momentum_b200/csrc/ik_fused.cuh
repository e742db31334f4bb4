// Fused Gauss-Newton solve: ONE persistent kernel runs SolverT::solve (solver.cpp:50-128) for the whole batch.
//
// A CTA holds up to three *instance groups* of 256 threads. Every table the iteration walks (character, plan, Gram plan, Cholesky
// schedule) is staged once per CTA by bulk copies; a group then takes instances off a device-side work counter and keeps each one
// in shared memory for ALL of its iterations:
//
//   theta --ParameterTransform / FK sweep--> joint frames --units/cells--> Jacobian strips (shared memory)
//         --tile-sparse Gram (mma.sync 3xTF32)--> accumulators parked in TMEM while the strips are still being read
//         --tcgen05.ld--> 16x16 tiles of J^T J + lambda I written over the (now dead) strips
//         --level-scheduled Cholesky + both substitutions--> delta --> theta -= delta, convergence test (solver.cpp:98-113)
//
// so per solve an instance moves theta, its targets and three result words through HBM and nothing else: no Jacobian, no normal
// matrix, no per-iteration launch, no host round trip for the stopping rule (an instance that converges frees its group at once).
// The device functions are the ones the multi-kernel path uses (ik_device.cuh, ik_chol_sched.cuh): same arithmetic, same order.
#pragma once

#include <cuda_runtime.h>

#include "ik_chol_sched.cuh"
#include "ik_types.h"

namespace mb2 {

constexpr int kFusedGroupThreads = 256;
constexpr int kFusedMaxGroups = 3;

// word (4-byte) offsets of the tables inside the plan blob (every table starts on a 16-byte boundary)
struct FusedBlobLayout {
  int32_t parent, offset, prerot, ptOuter, ptInner, ptVals, ptOffsets, levelStart, levelJoints;
  int32_t efs, units, cells, contribs, limitData;
  int32_t cols;       // device column -> model parameter (-1: alignment column)
  int32_t gram;       // Gram plan blob (makeGramBlob), gramOff[] are relative to it
  int32_t sched;      // Cholesky schedule blob (makeScheduleBlob)
  int32_t words;      // total size, multiple of 4
};

struct FusedArgs {
  int32_t batch;
  FunctionTables T;            // sizes; the table pointers are replaced by shared-memory addresses inside the kernel
  const int32_t* blob;         // the plan blob in global memory
  FusedBlobLayout L;
  CholSchedDev S;              // table pointers relative to `schedGlobal` (rebased onto the staged copy)
  const int32_t* schedGlobal;  // = S.blob
  int32_t gramOff[8];          // makeGramBlob offsets
  int32_t numStrips, stripStride, residOff, numOrder; // GramPlan (numOrder: entries of the tile-order table)
  float regularization, threshold;
  int32_t minIterations, maxIterations;
  float* theta;                // [B][ldTheta] in/out
  int32_t ldTheta;
  const float* targets;        // [B][T.targetStride]
  const float* cweights;       // [numWeights] or [B][numWeights]
  double* errors;              // [B] objective before the last update (what SolverT::solve returns)
  int32_t* iterations;         // [B]
  int32_t* status;             // [B] mb2_instance_status
  double* history;             // optional [B][maxIterations]
  int32_t* workCounter;        // zeroed before the launch
  unsigned long long* phaseCycles; // optional [16]: per-phase cycles of the first group of CTA 0 (profiling instantiation)
  int32_t groups;              // instance groups per CTA (1..3)
  int32_t groupFloats;         // shared-memory floats per group
  int32_t unionFloats;         // ... of which the strips / tiles union region
  int32_t tmemColsPerWarp;     // 8 * tiles per warp
};

struct FusedConfig {
  int groups{0};
  int groupFloats{0}, unionFloats{0}, tmemColsPerWarp{0};
  size_t smemBytes{0};
};
// how many groups fit next to the staged tables (0: the fused kernel cannot run this plan)
// orderRounds: rounds of the Gram tile-order table (tiles per warp at most)
FusedConfig fusedConfigure(const FunctionTables& T, const FusedBlobLayout& L, const CholSchedDev& S, int stripStride, int orderRounds, int maxSmemOptin);
cudaError_t launchFusedSolve(const FusedArgs& a, const FusedConfig& cfg, int numSms, bool profile, cudaStream_t stream);

} // namespace mb2

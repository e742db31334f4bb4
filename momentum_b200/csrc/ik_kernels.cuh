// sm_100a kernels of the batched Gauss-Newton iteration (declarations + launch parameter structs).
#pragma once

#include <cuda_runtime.h>

#include "ik_chol_sched.cuh"
#include "ik_types.h"

namespace mb2 {

struct SweepArgs {           // K1 (FK + residual + Jacobian) and K4 (FK + error only)
  FunctionTables T;
  int32_t batch;
  const float* theta;        // [B][ldTheta]
  int32_t ldTheta;
  const float* targets;      // [B][T.targetStride]
  const float* cweights;     // [numWeights] or [B][numWeights]
  float* jacobian;           // [B][numCols + 1][ldJ] (K1 only); column numCols is the residual vector
  double* errors;            // [B]
  const int32_t* active;     // optional per-instance mask
  float* stateOut;           // optional [B][J][8]
  int32_t stageTables;       // set by launchSweep: 1 every read-only table in shared memory, 2 all but cells / contributions, 0 none
  int32_t warpsPerInstance;  // set by launchSweep: 1, 2, 4 or 8 warps share one instance (large rigs: few instances fit in shared memory)
};

struct JtJArgs {             // K2
  int32_t batch;
  const float* jacobian;     // [B][numCols + 1][ldJ]; column numCols = residual
  int32_t numCols, ldJ, kRows; // kRows = contraction length (rows rounded up to 4)
  int32_t ns;                // leading ns columns enter the normal equations (ns <= numCols)
  float* H;                  // [B][ns+1][ldH] symmetric matrix [J r]^T [J r] restricted to the leading ns columns + r, row-major;
                             // only the UPPER triangle H[i*ldH + j], j >= i, is guaranteed (column ns = J^T r: H[i*ldH + ns] = (J^T r)_i)
  int32_t ldH;               // multiple of 16, >= ns + 1
  size_t hStride;            // floats per instance in H
  const int32_t* active;
  float* g;                  // optional [B][ldG]: J^T r again, as a contiguous vector
  int32_t ldG;
};

struct CholArgs {            // K3: damped Cholesky + solve + update + SolverT bookkeeping
  int32_t batch;
  float* H;                  // [B][ns+1][ldH] full symmetric [JtJ, Jtr] (as written by K2); K3 never writes it unless it factors in place
  size_t hStride;            // floats per instance in H
  int32_t ns, ldH;
  float regularization;
  const int32_t* cols;       // [ns] subset -> full parameter index
  float* theta;              // [B][ldTheta], updated: theta[cols[a]] -= delta[a]   (no line search)
  int32_t ldTheta;
  float* delta;              // [B][ns] (always written)
  int32_t applyUpdate;       // 1: theta -= delta here
  const double* errors;      // [B] error of this iteration (from K1)
  double* lastErrors;        // [B]
  int32_t* active;           // [B] in/out
  int32_t* iterations;       // [B]
  int32_t* status;           // [B]
  double* history;           // optional [B][maxIterations]
  int32_t iteration, minIterations, maxIterations;
  float threshold;
  int32_t* activeCount;      // device counter (atomicAdd of instances still active)
  int32_t bookkeeping;       // 1: run the SolverT convergence test here (no line search)
  float* gradDotDelta;       // optional [B]: Jtr . delta (SubsetGaussNewtonSolverT line search)
  const float* g;            // scheduled kernel: [B][ldG] J^T r as a contiguous vector (ldG = ns rounded up to 4)
  int32_t ldG;
  const float* tilesIn;      // scheduled kernel, optional: [B][tilesStride] tiles + slot-ordered J^T r from the Gram kernel (H, g unused)
  size_t tilesStride;
  int32_t profile;           // MB2_CHOL_PROFILE=1: block 0 prints per-phase cycles (debug aid)
};

struct GramArgs {            // K2s: stored tiles of J^T J + lambda I and J^T r from the non-zero strips of the Jacobian (GramPlan)
  int32_t batch;
  const float* strips;       // [B][stripStride]: the Jacobian in strip layout + residual (FunctionTables::stripMode), written by the sweep kernel
  size_t stripStride;        // GramPlan::stride
  int32_t residOff;          // GramPlan::residOff
  const int32_t* active;
  int32_t numStrips, numTiles, numTileCols, nPad;
  int32_t numOrder;          // entries of the tile-order table ([rounds][kGramWarps], -1 = idle)
  // the GramPlan tables as one int32 blob (staged in shared memory by the kernel); offsets in ints
  const int32_t* blob;
  int32_t blobInts;
  int32_t offTileOrder, offTilePairStart, offPairA, offPairB, offColStripStart, offColStrip, offStripRow, offTileInfo;
  float regularization;
  float* out;                // [B][outStride]: numTiles x 256 floats in tile storage order, then the slot-ordered J^T r [nPad]
  size_t outStride;
};
cudaError_t launchGramTiles(const GramArgs& a, cudaStream_t stream);
size_t gramTilesSmemBytes(size_t stripStride, int blobInts);

// K2s + K3 in one launch: strips in (bulk copy), Gram accumulators parked in TMEM, tiles written over the strips, tile Cholesky,
// update. No tile round trip through HBM and no second launch; `g.out` is unused.
struct GramCholArgs {
  GramArgs g;
  CholArgs c;
  unsigned long long* phaseCycles; // optional [8], profiling instantiation: prologue, gram, tiles from TMEM, diag, panel, update, backward, finish (block 0)
};
size_t gramCholeskySmemBytes(size_t stripStride, int gramBlobInts, int n, int nPad, int numTiles, int schedBlobInts);
cudaError_t launchGramCholesky(const GramCholArgs& a, const CholSchedDev& sched, bool profile, cudaStream_t stream);

// QR-accurate linear step (ik_qr.cu): Householder sweeps of the K-major Jacobian into a shared-memory R, replaces JtJ + Cholesky
struct QrArgs {
  CholArgs c;                // ns, regularization, cols, theta, delta, bookkeeping ... (H / tiles unused)
  const float* jacobian;     // [B][numCols + 1][ldJ], column numCols = residual (compact plan: numCols = ns)
  int32_t numCols, ldJ;
  const int32_t* chunkStart; // [numChunks + 1] first row of every chunk (device memory); a chunk never crosses an error-function block
  int32_t numChunks;
};
// TrustRegionQRT's iteration (ik_tr_qr.cuh): the QR arguments plus what getError needs inside the kernel and the solver's radius state
struct TrQrArgs {
  QrArgs q;
  FunctionTables T;
  const float* targets;
  const float* cweights;
  float* radius;      // [B] curTrustRegionRadius_ (in / out)
  float* rSaved;      // [B][packed upper triangle, rounded up to 4 floats] scratch for Rmatrix_
  float maxRadius;    // maxTrustRegionRadius_ = 10 (trust_region_qr.h:73)
  int32_t maxChunkRows;
};
size_t trQrSmemFloats(int n, int numParams, int numJoints, int maxChunkRows);
cudaError_t launchTrustRegionQr(const TrQrArgs& a, cudaStream_t stream);
size_t qrSmemFloats(int n, int maxChunkRows);
int qrMaxChunkRows(int n, size_t smemBytes); // rows of Jacobian that fit beside R (0: the system is too large for this kernel)
cudaError_t launchQrSolve(const QrArgs& a, int maxChunkRows, cudaStream_t stream);
cudaError_t launchSweep(const SweepArgs& a, bool jacobian, cudaStream_t stream);
size_t sweepSmemPerInstance(const FunctionTables& T, int warpsPerInstance);
cudaError_t launchJtJSimt(const JtJArgs& a, cudaStream_t stream);
cudaError_t launchCholesky(const CholArgs& a, cudaStream_t stream);
// level-scheduled tile-sparse variant (ik_chol_sched.h); returns cudaErrorInvalidConfiguration when the tiles do not fit in shared memory
// `a.H` is the full symmetric system in device-column (= elimination) order
cudaError_t launchCholeskyScheduled(const CholArgs& a, const CholSchedDev& sched, cudaStream_t stream);
inline int cholGradientLd(int ns) { return (ns + 3) & ~3; }
size_t choleskyScheduledSmemBytes(int n, int nPad, int numTiles, int blobInts);
cudaError_t initKernelAttributes();

// line-search helpers
cudaError_t launchTrialUpdate(int batch, const float* thetaOrig, int ldTheta, const float* delta, int ns, const int32_t* cols, const float* scale /*[B]*/,
                              float* thetaTrial, const int32_t* active, cudaStream_t stream);
struct LineSearchArgs {
  int32_t batch, ns, numParams, ldTheta;
  const double* errors;      // error_ at theta (from the Jacobian pass)
  const double* trialErrors; // getError(theta - scale*delta)
  const float* gradDotDelta; // [B] (subset variant) or nullptr
  float* scale;              // [B] in/out line-search scale
  int32_t* searching;        // [B] 1 while the instance is still halving
  int32_t step;              // 0..9
  int32_t subsetVariant;
  const int32_t* active;
};
cudaError_t launchLineSearchStep(const LineSearchArgs& a, cudaStream_t stream);
cudaError_t launchCommitTrial(int batch, int numParams, int ldTheta, const float* thetaTrial, float* theta, const int32_t* active, cudaStream_t stream);
struct BookkeepingArgs {
  int32_t batch;
  const double* errors; double* lastErrors; int32_t* active; int32_t* iterations; int32_t* status; double* history;
  const float* theta; int32_t ldTheta, numParams;
  int32_t iteration, minIterations, maxIterations; float threshold; int32_t* activeCount;
};
cudaError_t launchBookkeeping(const BookkeepingArgs& a, cudaStream_t stream);
cudaError_t launchScatterTargets(const float* packed, float* dst, int size, int strideFloats, int batch, cudaStream_t stream);
cudaError_t launchNormalizeQuats(float* base, int count, int strideFloats, int quatsPerRecord, int batch, cudaStream_t stream);

} // namespace mb2

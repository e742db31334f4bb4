// C-ABI of momentum_b200 (include/momentum_b200.h): handle management, host<->device plumbing and
// the batched SolverT::solve driver. No CPU fallback: every compute entry fails with MB2_ERR_CUDA
// when no sm_100 device is usable.
#include "../../include/momentum_b200.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ik_chol_sched.h"
#include "ik_fused.cuh"
#include "ik_jtj_tc.cuh"
#include "ik_kernels.cuh"
#include "ik_plan.h"

using namespace mb2;

namespace {

thread_local std::string g_lastError;

int fail(int code, const std::string& msg) {
  g_lastError = msg;
  return code;
}
#define MB2_CUDA(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) return fail(MB2_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)
#define MB2_CHECK(cond, msg)                                   \
  do {                                                         \
    if (!(cond)) return fail(MB2_ERR_INVALID_ARGUMENT, (msg)); \
  } while (0)

template <class T>
struct DeviceBuffer {
  T* p{nullptr};
  size_t n{0};
  ~DeviceBuffer() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t resize(size_t count) {
    if (count <= n && p) return cudaSuccess;
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = resize(std::max<size_t>(v.size(), 1));
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

int roundUp(int v, int m) { return (v + m - 1) / m * m; }

int usableDevices();

// Every entry point that allocates or launches runs with the handle's device current and restores the caller's device on exit
// (a single process may drive several GPUs: one character / solver function per device).
struct DeviceGuard {
  int prev{-1};
  bool switched{false};
  cudaError_t err{cudaSuccess};
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) { err = cudaSetDevice(device); switched = err == cudaSuccess; }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define MB2_DEVICE_GUARD(device)                                                                                      \
  if (usableDevices() <= 0) return fail(MB2_ERR_CUDA, "no usable sm_100 CUDA device: momentum_b200 has no CPU fallback"); \
  DeviceGuard _guard(device);                                                                                         \
  MB2_CUDA(_guard.err)

// NVTX ranges named like the reference's MT_PROFILE_FUNCTION zones (solver.cpp:51, gauss_newton_solver.cpp:225, skeleton_state.cpp:88)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

} // namespace

struct mb2_character {
  int device{0};
  HostCharacter host;
  DeviceBuffer<int32_t> parent, ptOuter, ptInner, levelStart, levelJoints;
  DeviceBuffer<float> offset, prerot, ptVals, ptOffsets;
  uint64_t limitsVersion{0};
};

struct DeviceSchedule {
  CholSchedule host;
  CholSchedDev dev{};
  bool valid{false};
  bool dense{false};
  DeviceBuffer<int32_t> blob;
  GramPlan gram;          // tile-sparse Gram tables of the same plan (rows aligned to quads)
  bool gramValid{false};
  DeviceBuffer<int32_t> gBlob;
  int32_t gBlobInts{0}, gOffsets[8]{};
  // fused persistent kernel (ik_fused.cuh): every table of the plan as one blob + how many instance groups fit beside it
  DeviceBuffer<int32_t> fBlob;
  FusedBlobLayout fLayout{};
  FusedConfig fConfig{};
};

struct mb2_solver_function {
  const mb2_character* ch{nullptr};
  int B{0};
  cudaStream_t stream{nullptr};
  std::vector<HostErrorFunction> efs;
  std::vector<uint8_t> enabled;
  int targetStride{0}, numWeights{0};
  bool weightsPerInstance{false};
  bool planDirty{true};
  int planMode{0};        // 0 full columns (API parity), 1 solver: enabled columns in natural order, 2 solver: elimination order + tile schedule
  bool planSchedDense{false};
  bool planAlignRows{false}; // row groups aligned to 4 (tile-sparse Gram reads the Jacobian in 4-row strips)
  uint64_t planLimitsVersion{~0ull};
  Plan plan;
  int ldJ{32};
  // device tables
  DeviceBuffer<EfDesc> dEfs;
  DeviceBuffer<UnitDesc> dUnits;
  DeviceBuffer<CellDesc> dCells;
  DeviceBuffer<ContribDesc> dContribs;
  DeviceBuffer<float> dLimitData;
  DeviceBuffer<int32_t> dEnabledList, dIdentity, dDeviceCols;
  // device data
  DeviceBuffer<float> dTargets, dWeights, dJ, dTheta, dState, dH, dTargetStage;
  DeviceBuffer<double> dErrors;
  std::vector<float> hWeights; // shared weights mirror
  std::unique_ptr<DeviceSchedule> sched; // Cholesky schedule of the current (compact) plan
  FunctionTables tables() const;
};

struct PhaseEvent {
  cudaEvent_t start, stop;
  int phase;
};

struct mb2_solver {
  mb2_solver_function* fn{nullptr};
  mb2_gauss_newton_options opt{};
  DeviceBuffer<float> dH, dDelta, dThetaOrig, dTheta0, dScale, dGradDotDelta, dThetaStage, dGrad, dTiles, dRadius, dRSaved;
  DeviceBuffer<double> dLastErrors, dTrialErrors, dHistory;
  DeviceBuffer<int32_t> dActive, dIterations, dStatus, dSearching, dActiveCount, dWorkCounter, dQrChunks;
  DeviceBuffer<unsigned long long> dPhaseCycles;
  bool lastFused{false}, lastGramChol{false};
  int lastFusedGroups{0};
  cudaEvent_t fusedStart{nullptr}, fusedStop{nullptr};
  double fusedMs{0};
  int* hActiveCount{nullptr}; // pinned
  // pinned staging of the per-instance results of an asynchronous host solve: [B] errors (double) | [B] iterations | [B] status, filled by
  // copies enqueued behind the solve so that mb2_solver_wait needs one stream synchronisation and no further device round trip
  char* hResults{nullptr};
  size_t hResultsBytes{0};
  bool resultsStaged{false};
  uint64_t totalIterations{0}, kernelLaunches{0};
  bool profiling{false};
  bool inKernelProfile{false}; // profiling level 2: the instrumented kernel instantiations (clock64 per phase; slower)
  std::vector<PhaseEvent> events;
  double phaseMs[4]{0, 0, 0, 0};
  uint64_t phaseLaunches[4]{0, 0, 0, 0};
  size_t historyStride{0};
  ~mb2_solver() {
    if (hActiveCount) cudaFreeHost(hActiveCount);
    if (hResults) cudaFreeHost(hResults);
    if (fusedStart) cudaEventDestroy(fusedStart);
    if (fusedStop) cudaEventDestroy(fusedStop);
    for (auto& e : events) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
  }
};

FunctionTables mb2_solver_function::tables() const {
  FunctionTables T{};
  const HostCharacter& h = ch->host;
  T.numJoints = h.numJoints;
  T.numParams = h.numParams;
  T.parent = ch->parent.p;
  T.offset = ch->offset.p;
  T.prerot = ch->prerot.p;
  T.ptOuter = ch->ptOuter.p;
  T.ptInner = ch->ptInner.p;
  T.ptVals = ch->ptVals.p;
  T.ptOffsets = ch->ptOffsets.p;
  T.numLevels = int(h.levelStart.size()) - 1;
  T.levelStart = ch->levelStart.p;
  T.levelJoints = ch->levelJoints.p;
  T.numEf = int(plan.efs.size());
  T.numUnits = int(plan.units.size());
  T.numCells = int(plan.cells.size());
  T.efs = dEfs.p;
  T.units = dUnits.p;
  T.cells = dCells.p;
  T.contribs = dContribs.p;
  T.limitData = dLimitData.p;
  T.targetStride = targetStride;
  T.recStride = plan.recStride;
  T.numRows = plan.numRows;
  T.ldJ = ldJ;
  T.numCols = plan.numCols;
  T.weightsPerInstance = weightsPerInstance ? 1 : 0;
  T.numWeights = numWeights;
  T.ptNnz = int(h.ptInner.size());
  T.numContribs = int(plan.contribs.size());
  T.numLimitData = int(plan.limitData.size());
  const bool strips = planMode == 2 && planAlignRows && sched && sched->gramValid;
  T.stripMode = strips ? 1 : 0;
  T.residOff = strips ? sched->gram.residOff : 0;
  T.jacobianStride = strips ? size_t(sched->gram.stride) : size_t(plan.numCols + 1) * ldJ;
  return T;
}

namespace {

bool g_deviceChecked = false;
int g_usableDevices = 0;
} // namespace
namespace {
int usableDevices() {
  if (!g_deviceChecked) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { n = 0; cudaGetLastError(); }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
      int major = 0;
      if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    g_usableDevices = ok;
    g_deviceChecked = true;
  }
  return g_usableDevices;
}

int requireDevice(int device) {
  if (usableDevices() <= 0) return fail(MB2_ERR_CUDA, "no usable sm_100 CUDA device: momentum_b200 has no CPU fallback");
  MB2_CUDA(cudaSetDevice(device));
  MB2_CUDA(initKernelAttributes());
  return MB2_OK;
}

int ensureTargets(mb2_solver_function* f) {
  const size_t need = size_t(f->B) * std::max(f->targetStride, 1);
  if (f->dTargets.n < need) {
    DeviceBuffer<float> old;
    std::swap(old.p, f->dTargets.p);
    std::swap(old.n, f->dTargets.n);
    MB2_CUDA(f->dTargets.resize(need));
    MB2_CUDA(cudaMemsetAsync(f->dTargets.p, 0, need * sizeof(float), f->stream));
    // the next writer may be mb2_set_targets_device on a caller's stream: the zero fill must have landed first
    MB2_CUDA(cudaStreamSynchronize(f->stream));
    (void)old; // targets of blocks added earlier must be re-sent after adding blocks (documented)
  }
  return MB2_OK;
}

int uploadWeights(mb2_solver_function* f) {
  if (!f->weightsPerInstance) {
    std::vector<float> w = f->hWeights;
    if (w.empty()) w.push_back(0.f);
    MB2_CUDA(f->dWeights.upload(w, f->stream));
  }
  return MB2_OK;
}

// depth (in the joint tree) of the deepest joint each enabled parameter drives: tie-break priority of the elimination order
static std::vector<int> columnDepthPriority(const HostCharacter& h, const std::vector<int32_t>& enabledList) {
  std::vector<int> jointDepth(h.numJoints, 0);
  for (int j = 0; j < h.numJoints; ++j) { int d = 0; for (int a = h.parent[j]; a >= 0; a = h.parent[a]) ++d; jointDepth[j] = d; }
  std::vector<int> paramDepth(h.numParams, 0);
  for (int r = 0; r < kParametersPerJoint * h.numJoints; ++r)
    for (int k = h.ptOuter[r]; k < h.ptOuter[r + 1]; ++k) paramDepth[h.ptInner[k]] = std::max(paramDepth[h.ptInner[k]], jointDepth[r / kParametersPerJoint]);
  std::vector<int> prio(enabledList.size());
  for (size_t a = 0; a < enabledList.size(); ++a) prio[a] = paramDepth[enabledList[a]];
  return prio;
}

int uploadSchedule(mb2_solver_function* f, std::unique_ptr<DeviceSchedule>& ds) {
  std::vector<int32_t> blob;
  CholSchedDev hostView;
  makeScheduleBlob(ds->host, blob, hostView);
  MB2_CUDA(ds->blob.upload(blob, f->stream));
  ds->dev = rebaseSchedule(hostView, ds->blob.p);
  ds->valid = true;
  return MB2_OK;
}

// Every table the fused kernel walks, as one 16-byte-aligned blob (bulk-copied into shared memory once per CTA).
int buildFusedBlob(mb2_solver_function* f, DeviceSchedule& ds, const std::vector<int32_t>& gramBlob, const std::vector<int32_t>& schedBlob, cudaStream_t s) {
  std::vector<int32_t> blob;
  auto add = [&](const void* data, size_t bytes) {
    const int32_t off = int32_t(blob.size());
    const size_t words = (bytes + 3) / 4;
    blob.resize(blob.size() + words, 0);
    if (bytes) std::memcpy(blob.data() + off, data, bytes);
    while (blob.size() % 4) blob.push_back(0);
    return off;
  };
  const HostCharacter& h = f->ch->host;
  const Plan& p = f->plan;
  FusedBlobLayout& L = ds.fLayout;
  L.parent = add(h.parent.data(), h.parent.size() * 4);
  L.offset = add(h.offset.data(), h.offset.size() * 4);
  L.prerot = add(h.prerot.data(), h.prerot.size() * 4);
  L.ptOuter = add(h.ptOuter.data(), h.ptOuter.size() * 4);
  L.ptInner = add(h.ptInner.data(), h.ptInner.size() * 4);
  L.ptVals = add(h.ptVals.data(), h.ptVals.size() * 4);
  L.ptOffsets = add(h.ptOffsets.data(), h.ptOffsets.size() * 4);
  L.levelStart = add(h.levelStart.data(), h.levelStart.size() * 4);
  L.levelJoints = add(h.levelJoints.data(), h.levelJoints.size() * 4);
  L.efs = add(p.efs.data(), p.efs.size() * sizeof(EfDesc));
  L.units = add(p.units.data(), p.units.size() * sizeof(UnitDesc));
  L.cells = add(p.cells.data(), p.cells.size() * sizeof(CellDesc));
  L.contribs = add(p.contribs.data(), p.contribs.size() * sizeof(ContribDesc));
  L.limitData = add(p.limitData.data(), p.limitData.size() * 4);
  L.cols = add(p.deviceCols.data(), p.deviceCols.size() * 4);
  L.gram = add(gramBlob.data(), gramBlob.size() * 4);
  L.sched = add(schedBlob.data(), schedBlob.size() * 4);
  if (blob.empty()) blob.push_back(0);
  while (blob.size() % 4) blob.push_back(0);
  L.words = int32_t(blob.size());
  MB2_CUDA(ds.fBlob.upload(blob, s));
  return MB2_OK;
}

// mode 0: every column at its model-parameter index (getJacobian / getJtJR parity);
// mode 1: solver, only the enabled columns, ascending (dense Eigen-structured Cholesky);
// mode 2: solver, enabled columns in the Cholesky elimination order + tile schedule (schedDense: dense pattern).
// Modes 0 and 1 coincide when every parameter is enabled.
int ensurePlan(mb2_solver_function* f, int mode, bool schedDense = false, bool alignRows = false) {
  bool allEnabled = true;
  for (uint8_t e : f->enabled) allEnabled = allEnabled && e;
  if (allEnabled && mode == 1) mode = 0;
  if (mode != 2) alignRows = false;
  if (!f->planDirty && f->planLimitsVersion == f->ch->limitsVersion && f->planMode == mode && (mode != 2 || (f->planSchedDense == schedDense && f->planAlignRows == alignRows)))
    return MB2_OK;
  MB2_DEVICE_GUARD(f->ch->device);
  cudaStream_t s = f->stream;
  f->sched.reset();
  std::string err = buildPlan(f->ch->host, f->efs, f->enabled, mode != 0, f->plan);
  if (!err.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, err);
  if (mode == 2) {
    auto ds = std::make_unique<DeviceSchedule>();
    const int ns = f->plan.numCols;
    std::vector<std::vector<int>> cliques(f->plan.units.size());
    for (const CellDesc& c : f->plan.cells) cliques[c.unit].push_back(int(c.col));
    const std::vector<int> prio = columnDepthPriority(f->ch->host, f->plan.enabledList);
    err = buildCholSchedule(ns, cliques, schedDense, ds->host, &prio);
    if (!err.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, err);
    // re-plan with the device columns in elimination order (tile starts aligned to 4 columns), the schedule expressed in them
    std::vector<int32_t> colOrder;
    layoutDeviceColumns(ds->host, colOrder);
    for (int32_t& c : colOrder) if (c >= 0) c = f->plan.enabledList[c];
    err = buildPlan(f->ch->host, f->efs, f->enabled, true, f->plan, &colOrder, alignRows);
    if (!err.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, err);
    if (alignRows) {
      std::vector<int32_t> cr0, crn, cc;
      for (const CellDesc& c : f->plan.cells) { cr0.push_back(f->plan.units[c.unit].row0); crn.push_back(f->plan.units[c.unit].numRows); cc.push_back(int32_t(c.col)); }
      err = buildGramPlan(ds->host, cr0, crn, cc, f->plan.numRows, ds->gram);
      if (!err.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, err);
      for (size_t i = 0; i < f->plan.cells.size(); ++i) { f->plan.cells[i].stripOff = ds->gram.cellStripOff[i]; f->plan.cells[i].quadStride = ds->gram.cellQuadStride[i]; }
      std::vector<int32_t> gblob;
      makeGramBlob(ds->gram, ds->host, gblob, ds->gOffsets);
      ds->gBlobInts = int32_t(gblob.size());
      MB2_CUDA(ds->gBlob.upload(gblob, s));
      ds->gramValid = true;
    }
    ds->dense = schedDense;
    int rc = uploadSchedule(f, ds);
    if (rc != MB2_OK) return rc;
    if (alignRows) {
      std::vector<int32_t> gblob2, sblob;
      int32_t goff[8];
      makeGramBlob(ds->gram, ds->host, gblob2, goff);
      CholSchedDev hostView;
      makeScheduleBlob(ds->host, sblob, hostView);
      rc = buildFusedBlob(f, *ds, gblob2, sblob, s);
      if (rc != MB2_OK) return rc;
    }
    f->sched = std::move(ds);
  }
  f->planMode = mode;
  f->planSchedDense = schedDense;
  f->planAlignRows = alignRows;
  MB2_CUDA(f->dEfs.upload(f->plan.efs, s));
  MB2_CUDA(f->dUnits.upload(f->plan.units, s));
  MB2_CUDA(f->dCells.upload(f->plan.cells, s));
  MB2_CUDA(f->dContribs.upload(f->plan.contribs, s));
  MB2_CUDA(f->dLimitData.upload(f->plan.limitData, s));
  MB2_CUDA(f->dEnabledList.upload(f->plan.enabledList, s));
  MB2_CUDA(f->dDeviceCols.upload(f->plan.deviceCols, s));
  std::vector<int32_t> ident(f->ch->host.numParams);
  for (size_t i = 0; i < ident.size(); ++i) ident[i] = int32_t(i);
  MB2_CUDA(f->dIdentity.upload(ident, s));
  f->ldJ = std::max(32, roundUp(f->plan.numRows, 32));
  const bool stripLayout = mode == 2 && alignRows; // strips + residual instead of the K-major matrix
  const size_t jElems = stripLayout ? size_t(f->B) * size_t(f->sched->gram.stride) : size_t(f->B) * (f->plan.numCols + 1) * f->ldJ;
  MB2_CUDA(f->dJ.resize(jElems));
  // cells outside the plan are never written: zero once per plan (ResizeableMatrix::resizeAndSetZero
  // happens every iteration in the reference, solver_function.cpp:96)
  MB2_CUDA(cudaMemsetAsync(f->dJ.p, 0, jElems * sizeof(float), s));
  MB2_CUDA(f->dErrors.resize(f->B));
  MB2_CUDA(f->dTheta.resize(size_t(f->B) * f->ch->host.numParams));
  int rc = ensureTargets(f);
  if (rc != MB2_OK) return rc;
  rc = uploadWeights(f);
  if (rc != MB2_OK) return rc;
  if (stripLayout && f->sched) {
    int dev = 0, optin = 0;
    MB2_CUDA(cudaGetDevice(&dev));
    MB2_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    f->planMode = mode; // tables() reads the mode
    f->planAlignRows = alignRows;
    f->sched->fConfig = fusedConfigure(f->tables(), f->sched->fLayout, f->sched->dev, f->sched->gram.stride, int(f->sched->gram.tileOrder.size()) / kGramWarps, optin);
  }
  // Plan tables, schedule / Gram blobs, weights and the Jacobian zero fill were issued on the handle's stream; the solve may run on a
  // caller's stream. A plan build is a rare host-side event: wait for it here instead of threading events through every launch.
  MB2_CUDA(cudaStreamSynchronize(s));
  f->planDirty = false;
  f->planLimitsVersion = f->ch->limitsVersion;
  return MB2_OK;
}

SweepArgs sweepArgs(mb2_solver_function* f, const float* theta, const int32_t* active) {
  SweepArgs a{};
  a.T = f->tables();
  a.batch = f->B;
  a.theta = theta;
  a.ldTheta = f->ch->host.numParams;
  a.targets = f->dTargets.p;
  a.cweights = f->dWeights.p;
  a.jacobian = f->dJ.p;
  a.errors = f->dErrors.p;
  a.active = active;
  a.stateOut = nullptr;
  return a;
}

int addBlock(mb2_solver_function* f, HostErrorFunction& ef, int32_t* outIndex) {
  const bool hasWeights = ef.kind <= 2 || ef.kind == 5;
  if (hasWeights && f->weightsPerInstance) return fail(MB2_ERR_UNSUPPORTED, "add all error functions before setting per-instance constraint weights");
  ef.targetOff = f->targetStride;
  ef.weightOff = f->numWeights;
  f->targetStride += ef.targetSize;
  if (hasWeights) {
    f->numWeights += ef.numConstraints();
    f->hWeights.insert(f->hWeights.end(), ef.weights.begin(), ef.weights.end());
  }
  f->efs.push_back(ef);
  f->planDirty = true;
  if (outIndex) *outIndex = int32_t(f->efs.size()) - 1;
  return MB2_OK;
}

void bitsToEnabled(const uint64_t* bits, int n, std::vector<uint8_t>& out) {
  out.assign(n, 0);
  for (int i = 0; i < n; ++i) out[i] = (bits[i >> 6] >> (i & 63)) & 1ull ? 1 : 0;
}

void recordPhaseStart(mb2_solver* s, int phase, cudaStream_t st) {
  s->kernelLaunches++;
  if (!s->profiling) return;
  PhaseEvent e;
  cudaEventCreate(&e.start);
  cudaEventCreate(&e.stop);
  e.phase = phase;
  cudaEventRecord(e.start, st);
  s->events.push_back(e);
}
void recordPhaseStop(mb2_solver* s, cudaStream_t st) {
  if (!s->profiling) return;
  cudaEventRecord(s->events.back().stop, st);
}

__global__ void initSolveStateKernel(int batch, int32_t* active, int32_t* iterations, int32_t* status, double* lastErrors, double* errors) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  active[b] = 1;
  iterations[b] = 0;
  status[b] = 0;
  lastErrors[b] = DBL_MAX; // solver.cpp:83-84
  errors[b] = DBL_MAX;
}
__global__ void initLineSearchKernel(int batch, const int32_t* active, int32_t* searching, float* scale) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  searching[b] = active[b];
  scale[b] = 1.f;
}
// NaN/Inf guard of the batched caller (pymomentum/tensor_ik/tensor_ik.cpp:168-173): revert to the initial guess
__global__ void finalizeKernel(int batch, int n, float* theta, const float* theta0, int32_t* status) {
  const int b = blockIdx.x;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (!isfinite(theta[size_t(b) * n + i])) bad = 1;
  __syncthreads();
  if (bad) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) theta[size_t(b) * n + i] = theta0[size_t(b) * n + i];
    if (threadIdx.x == 0) status[b] = MB2_INSTANCE_NON_FINITE;
  }
}

int resolveJtjMode(const mb2_solver_function* f, int requested, int ns) {
  if (requested == MB2_JTJ_FP32_SIMT) return MB2_JTJ_FP32_SIMT;
  const bool ok = jtjTensorSupported(ns, f->plan.numCols, f->ldJ);
  if (requested == MB2_JTJ_AUTO) return ok ? MB2_JTJ_TF32X3 : MB2_JTJ_FP32_SIMT;
  return ok ? requested : -1;
}

int runJtJ(mb2_solver_function* f, int mode, int ns, float* H, int ldH, size_t hStride, const int32_t* active, cudaStream_t st, float* g = nullptr, int ldG = 0) {
  JtJArgs a{};
  a.batch = f->B;
  a.jacobian = f->dJ.p;
  a.numCols = f->plan.numCols;
  a.ldJ = f->ldJ;
  a.kRows = roundUp(std::max(f->plan.numRows, 1), 4);
  a.ns = ns;
  a.H = H;
  a.ldH = ldH;
  a.hStride = hStride;
  a.active = active;
  a.g = g;
  a.ldG = ldG;
  if (mode == MB2_JTJ_FP32_SIMT) { MB2_CUDA(launchJtJSimt(a, st)); }
  else { MB2_CUDA(launchJtJTensor(a, mode == MB2_JTJ_TF32X3 ? 3 : 1, st)); }
  return MB2_OK;
}

} // namespace

extern "C" {

const char* mb2_last_error(void) { return g_lastError.c_str(); }
int mb2_device_count(void) { return usableDevices(); }

void mb2_default_gauss_newton_options(mb2_gauss_newton_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->min_iterations = 1;  // solver.h:21
  o->max_iterations = 2;  // solver.h:24
  o->threshold = 1.0f;    // solver.h:27
  o->verbose = 0;
  o->regularization = 0.05f; // gauss_newton_solver.h:22
  o->do_line_search = 0;
  o->use_block_jtj = 0;
  o->target_rows_per_chunk = ~0ull;
  o->subset_line_search = 0;
  o->jtj_mode = MB2_JTJ_AUTO;
  o->store_error_history = 0;
  o->cholesky_mode = MB2_CHOLESKY_AUTO;
  o->fused_mode = MB2_FUSED_AUTO;
  o->linear_solver = MB2_LINEAR_SOLVER_CHOLESKY;
  o->trust_region_radius = 1.0f; // trust_region_qr.h:23
}

int mb2_character_create(int device, int32_t J, const int32_t* parents, const float* offsets, const float* prerot, int32_t n,
                         const int32_t* outer, const int32_t* inner, const float* vals, const float* ptoffsets, mb2_character** out) {
  MB2_CHECK(out != nullptr && parents && offsets && prerot && outer && ptoffsets, "null argument");
  MB2_CHECK(J > 0 && n > 0, "numJoints and numModelParameters must be positive");
  auto c = std::make_unique<mb2_character>();
  HostCharacter& h = c->host;
  h.numJoints = J;
  h.numParams = n;
  h.parent.assign(parents, parents + J);
  h.offset.assign(offsets, offsets + 3 * J);
  h.prerot.assign(prerot, prerot + 4 * J);
  h.ptOuter.assign(outer, outer + 7 * J + 1);
  const int nnz = outer[7 * J];
  MB2_CHECK(nnz >= 0 && (nnz == 0 || (inner && vals)), "parameter transform nnz invalid");
  h.ptInner.assign(inner, inner + nnz);
  h.ptVals.assign(vals, vals + nnz);
  h.ptOffsets.assign(ptoffsets, ptoffsets + 7 * J);
  const std::string err = h.validate();
  if (!err.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, err);
  h.buildLevels();
  MB2_DEVICE_GUARD(device);
  int rc = requireDevice(device);
  if (rc != MB2_OK) return rc;
  c->device = device;
  MB2_CUDA(c->parent.upload(h.parent, nullptr));
  MB2_CUDA(c->offset.upload(h.offset, nullptr));
  MB2_CUDA(c->prerot.upload(h.prerot, nullptr));
  MB2_CUDA(c->ptOuter.upload(h.ptOuter, nullptr));
  MB2_CUDA(c->ptInner.upload(h.ptInner, nullptr));
  MB2_CUDA(c->ptVals.upload(h.ptVals, nullptr));
  MB2_CUDA(c->ptOffsets.upload(h.ptOffsets, nullptr));
  MB2_CUDA(c->levelStart.upload(h.levelStart, nullptr));
  MB2_CUDA(c->levelJoints.upload(h.levelJoints, nullptr));
  MB2_CUDA(cudaStreamSynchronize(nullptr));
  *out = c.release();
  return MB2_OK;
}

int mb2_character_set_parameter_limits(mb2_character* c, int32_t count, const mb2_parameter_limit* limits) {
  MB2_CHECK(c != nullptr && count >= 0 && (count == 0 || limits), "invalid limits");
  c->host.limits.clear();
  for (int i = 0; i < count; ++i) {
    HostLimit l;
    l.type = limits[i].type;
    l.weight = limits[i].weight;
    std::memcpy(l.i, limits[i].i, sizeof(l.i));
    std::memcpy(l.f, limits[i].f, sizeof(l.f));
    MB2_CHECK(l.type >= 0 && l.type <= 6, "Unknown parameter type for joint limit");
    c->host.limits.push_back(l);
  }
  c->limitsVersion++;
  return MB2_OK;
}

void mb2_character_destroy(mb2_character* c) { delete c; }

int mb2_solver_function_create(const mb2_character* c, int32_t batch, mb2_solver_function** out) {
  MB2_CHECK(c != nullptr && out != nullptr, "null argument");
  MB2_CHECK(batch > 0, "batch must be positive");
  MB2_DEVICE_GUARD(c->device);
  int rc = requireDevice(c->device);
  if (rc != MB2_OK) return rc;
  auto f = std::make_unique<mb2_solver_function>();
  f->ch = c;
  f->B = batch;
  f->enabled.assign(c->host.numParams, 1); // all parameters enabled by default (skeleton_error_function.h:27)
  MB2_CUDA(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  *out = f.release();
  return MB2_OK;
}

void mb2_solver_function_destroy(mb2_solver_function* f) {
  if (!f) return;
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
}

// ---- replicas on other devices (single-process multi-GPU: ik_sharded.cpp) ----
// The rig (skeleton, parameter transform, parameter limits) on another device; independent of the original afterwards.
int mb2_character_clone(const mb2_character* c, int device, mb2_character** out) {
  MB2_CHECK(c != nullptr && out != nullptr, "null argument");
  const HostCharacter& h = c->host;
  mb2_character* copy = nullptr;
  int rc = mb2_character_create(device, h.numJoints, h.parent.data(), h.offset.data(), h.prerot.data(), h.numParams, h.ptOuter.data(), h.ptInner.data(), h.ptVals.data(),
                                h.ptOffsets.data(), &copy);
  if (rc != MB2_OK) return rc;
  copy->host.limits = h.limits;
  copy->limitsVersion = 1;
  *out = copy;
  return MB2_OK;
}
// The DEFINITION of a solver function (error-function blocks with their shared constraint data and weights, block weights, enabled
// parameters) for `batch` instances of character `c` (normally a clone of f's character on another device). Per-instance data
// (targets, per-instance weights / offsets) is not copied: it belongs to the instances the new function will hold.
int mb2_solver_function_clone(const mb2_solver_function* f, const mb2_character* c, int32_t batch, mb2_solver_function** out) {
  MB2_CHECK(f != nullptr && c != nullptr && out != nullptr, "null argument");
  MB2_CHECK(c->host.numJoints == f->ch->host.numJoints && c->host.numParams == f->ch->host.numParams, "clone target character has a different shape");
  MB2_CHECK(!f->weightsPerInstance, "a function with per-instance constraint weights cannot be cloned (set them on the clone)");
  mb2_solver_function* g = nullptr;
  int rc = mb2_solver_function_create(c, batch, &g);
  if (rc != MB2_OK) return rc;
  g->efs = f->efs;
  g->enabled = f->enabled;
  g->targetStride = f->targetStride;
  g->numWeights = f->numWeights;
  g->hWeights = f->hWeights;
  g->planDirty = true;
  *out = g;
  return MB2_OK;
}
int mb2_character_device(const mb2_character* c) { return c ? c->device : -1; }
const mb2_character* mb2_solver_function_character(const mb2_solver_function* f) { return f ? f->ch : nullptr; }
int32_t mb2_solver_function_num_error_functions(const mb2_solver_function* f) { return f ? int32_t(f->efs.size()) : 0; }
int32_t mb2_solver_function_target_size(const mb2_solver_function* f, int32_t index) {
  return (f && index >= 0 && index < int32_t(f->efs.size())) ? f->efs[index].targetSize : -1;
}

int32_t mb2_solver_function_num_parameters(const mb2_solver_function* f) { return f ? f->ch->host.numParams : 0; }
int32_t mb2_solver_function_batch(const mb2_solver_function* f) { return f ? f->B : 0; }
int32_t mb2_solver_function_actual_parameters(const mb2_solver_function* f) {
  if (!f) return 0;
  int ap = 0;
  for (int i = 0; i < f->ch->host.numParams; ++i) if (f->enabled[i]) ap = i + 1;
  return ap;
}
int32_t mb2_solver_function_jacobian_rows(const mb2_solver_function* f) {
  if (!f) return 0;
  int total = 0;
  for (const auto& ef : f->efs) if (ef.weight > 0.f) total += jacobianBlockSize(f->ch->host, ef);
  return roundUp(total, 8); // padToSimdAlignment (solver_function.h:27-29)
}
int32_t mb2_solver_function_jacobian_stride(const mb2_solver_function* f) {
  if (!f) return 0;
  int total = 0;
  for (const auto& ef : f->efs) if (ef.weight > 0.f) total += jacobianBlockSize(f->ch->host, ef);
  return std::max(32, roundUp(total, 32));
}

int mb2_add_position_error_function(mb2_solver_function* f, float weight, float alpha, float c, int32_t nc, const int32_t* parents,
                                    const float* offsets, const float* weights, int32_t* outIndex) {
  MB2_CHECK(f != nullptr && nc >= 0 && (nc == 0 || (parents && offsets && weights)), "invalid position constraints");
  MB2_CHECK(c > 0.f, "Parameter c should be positive"); // generalized_loss.cpp:83
  HostErrorFunction ef;
  ef.kind = 0;
  ef.weight = weight;
  ef.lossAlpha = alpha;
  ef.lossC = c;
  ef.parents.assign(parents, parents + nc);
  for (int p : ef.parents) MB2_CHECK(p >= 0 && p < f->ch->host.numJoints, "constraint parent joint out of range");
  ef.offsets.assign(offsets, offsets + 3 * size_t(nc));
  ef.weights.assign(weights, weights + nc);
  ef.targetSize = 3 * nc;
  return addBlock(f, ef, outIndex);
}

int mb2_add_position_error_function_instanced(mb2_solver_function* f, float weight, float alpha, float c, int32_t nc, const int32_t* parents, const float* weights,
                                              int32_t* outIndex) {
  MB2_CHECK(f != nullptr && nc >= 0 && (nc == 0 || (parents && weights)), "invalid position constraints");
  MB2_CHECK(c > 0.f, "Parameter c should be positive");
  HostErrorFunction ef;
  ef.kind = 0;
  ef.weight = weight;
  ef.lossAlpha = alpha;
  ef.lossC = c;
  ef.instanceOffsets = true;
  ef.parents.assign(parents, parents + nc);
  for (int p : ef.parents) MB2_CHECK(p >= 0 && p < f->ch->host.numJoints, "constraint parent joint out of range");
  ef.offsets.assign(3 * size_t(nc), 0.f);
  ef.weights.assign(weights, weights + nc);
  ef.targetSize = 6 * nc;
  return addBlock(f, ef, outIndex);
}

int mb2_add_plane_error_function(mb2_solver_function* f, float weight, float alpha, float c, int32_t above, int32_t nc, const int32_t* parents,
                                 const float* offsets, const float* weights, int32_t* outIndex) {
  MB2_CHECK(f != nullptr && nc >= 0 && (nc == 0 || (parents && offsets && weights)), "invalid plane constraints");
  MB2_CHECK(c > 0.f, "Parameter c should be positive");
  HostErrorFunction ef;
  ef.kind = 5;
  ef.weight = weight;
  ef.lossAlpha = alpha;
  ef.lossC = c;
  ef.halfPlane = above != 0;
  ef.parents.assign(parents, parents + nc);
  for (int p : ef.parents) MB2_CHECK(p >= 0 && p < f->ch->host.numJoints, "constraint parent joint out of range");
  ef.offsets.assign(offsets, offsets + 3 * size_t(nc));
  ef.weights.assign(weights, weights + nc);
  ef.targetSize = 4 * nc;
  return addBlock(f, ef, outIndex);
}

int mb2_add_model_parameters_error_function(mb2_solver_function* f, float weight, const float* targetWeights, int32_t* outIndex) {
  MB2_CHECK(f != nullptr && targetWeights != nullptr, "invalid model-parameter weights");
  HostErrorFunction ef;
  ef.kind = 6;
  ef.weight = weight;
  ef.paramWeights.assign(targetWeights, targetWeights + f->ch->host.numParams);
  ef.targetSize = f->ch->host.numParams;
  return addBlock(f, ef, outIndex);
}

int mb2_add_orientation_error_function(mb2_solver_function* f, float weight, float alpha, float c, int32_t rotDiff, int32_t nc,
                                       const int32_t* parents, const float* offsets, const float* weights, int32_t* outIndex) {
  MB2_CHECK(f != nullptr && nc >= 0 && (nc == 0 || (parents && offsets && weights)), "invalid orientation constraints");
  MB2_CHECK(c > 0.f, "Parameter c should be positive");
  HostErrorFunction ef;
  ef.kind = rotDiff ? 2 : 1;
  ef.weight = weight;
  ef.lossAlpha = alpha;
  ef.lossC = c;
  ef.parents.assign(parents, parents + nc);
  for (int p : ef.parents) MB2_CHECK(p >= 0 && p < f->ch->host.numJoints, "constraint parent joint out of range");
  ef.offsets.assign(offsets, offsets + 4 * size_t(nc));
  for (int i = 0; i < nc; ++i) { // OrientationDataT ctor: offset(inOffset.normalized()) (orientation_error_function.h:33-35)
    float* q = &ef.offsets[4 * size_t(i)];
    const float nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) q[k] /= nrm;
  }
  ef.weights.assign(weights, weights + nc);
  ef.targetSize = 4 * nc;
  return addBlock(f, ef, outIndex);
}

int mb2_add_state_error_function(mb2_solver_function* f, float weight, int32_t rotationErrorType, float posWgt, float rotWgt,
                                 const float* posW, const float* rotW, int32_t* outIndex) {
  MB2_CHECK(f != nullptr && posW && rotW, "invalid state error function");
  MB2_CHECK(rotationErrorType == 0 || rotationErrorType == 1, "unknown rotation error type");
  HostErrorFunction ef;
  ef.kind = 3;
  ef.weight = weight;
  ef.rotationErrorType = rotationErrorType;
  ef.posWgt = posWgt;
  ef.rotWgt = rotWgt;
  const int J = f->ch->host.numJoints;
  ef.posW.assign(posW, posW + J);
  ef.rotW.assign(rotW, rotW + J);
  ef.targetSize = 8 * J;
  return addBlock(f, ef, outIndex);
}

int mb2_add_limit_error_function(mb2_solver_function* f, float weight, float alpha, float c, int32_t* outIndex) {
  MB2_CHECK(f != nullptr, "null solver function");
  MB2_CHECK(c > 0.f, "Parameter c should be positive");
  HostErrorFunction ef;
  ef.kind = 4;
  ef.weight = weight;
  ef.lossAlpha = alpha;
  ef.lossC = c;
  ef.targetSize = 0;
  return addBlock(f, ef, outIndex);
}

int mb2_set_error_function_weight(mb2_solver_function* f, int32_t index, float weight) {
  MB2_CHECK(f != nullptr && index >= 0 && index < int(f->efs.size()), "error function index out of range");
  f->efs[index].weight = weight;
  f->planDirty = true;
  return MB2_OK;
}

static int setTargetsImpl(mb2_solver_function* f, int32_t index, const float* targets, bool deviceSrc, cudaStream_t st) {
  MB2_CHECK(f != nullptr && index >= 0 && index < int(f->efs.size()) && targets, "invalid targets");
  const HostErrorFunction& ef = f->efs[index];
  if (ef.targetSize == 0 && ef.kind != 4) return MB2_OK; // a block without constraints: nothing to send (setConstraints({}) in the reference)
  MB2_CHECK(ef.targetSize > 0, "this error function has no per-instance targets");
  MB2_DEVICE_GUARD(f->ch->device);
  int rc = ensureTargets(f);
  if (rc != MB2_OK) return rc;
  float* dst = f->dTargets.p + ef.targetOff;
  const float* packed = targets;
  if (!deviceSrc) { // one contiguous transfer, then a kernel spreads it into the per-instance records
    MB2_CUDA(f->dTargetStage.resize(size_t(f->B) * ef.targetSize));
    MB2_CUDA(cudaMemcpyAsync(f->dTargetStage.p, targets, size_t(f->B) * ef.targetSize * sizeof(float), cudaMemcpyHostToDevice, st));
    packed = f->dTargetStage.p;
  }
  MB2_CUDA(launchScatterTargets(packed, dst, ef.targetSize, f->targetStride, f->B, st));
  if (ef.kind == 1 || ef.kind == 2) MB2_CUDA(launchNormalizeQuats(dst, 0, f->targetStride, ef.numConstraints(), f->B, st));
  return MB2_OK;
}
int mb2_set_targets(mb2_solver_function* f, int32_t index, const float* targets) {
  if (!f) return fail(MB2_ERR_INVALID_ARGUMENT, "null solver function");
  int rc = setTargetsImpl(f, index, targets, false, f->stream);
  if (rc != MB2_OK) return rc;
  MB2_CUDA(cudaStreamSynchronize(f->stream)); // host buffer may be reused by the caller
  return MB2_OK;
}
int mb2_set_targets_device(mb2_solver_function* f, int32_t index, const float* targets, void* stream) {
  if (!f) return fail(MB2_ERR_INVALID_ARGUMENT, "null solver function");
  return setTargetsImpl(f, index, targets, true, stream ? (cudaStream_t)stream : f->stream);
}

int mb2_set_constraint_weights(mb2_solver_function* f, int32_t index, const float* weights, int32_t perInstance) {
  MB2_CHECK(f != nullptr && index >= 0 && index < int(f->efs.size()) && weights, "invalid constraint weights");
  HostErrorFunction& ef = f->efs[index];
  MB2_CHECK(ef.kind <= 2 || ef.kind == 5, "constraint weights apply to Position/Orientation/Plane error functions");
  MB2_DEVICE_GUARD(f->ch->device);
  const int nc = ef.numConstraints();
  if (!perInstance) {
    MB2_CHECK(!f->weightsPerInstance, "solver function already uses per-instance constraint weights");
    ef.weights.assign(weights, weights + nc);
    std::copy(weights, weights + nc, f->hWeights.begin() + ef.weightOff);
    return uploadWeights(f);
  }
  if (!f->weightsPerInstance) { // expand the shared array to [B][numWeights]
    std::vector<float> all(size_t(f->B) * f->numWeights);
    for (int b = 0; b < f->B; ++b) std::copy(f->hWeights.begin(), f->hWeights.end(), all.begin() + size_t(b) * f->numWeights);
    f->dWeights.release();
    MB2_CUDA(f->dWeights.upload(all, f->stream));
    f->weightsPerInstance = true;
  }
  MB2_CUDA(cudaMemcpy2DAsync(f->dWeights.p + ef.weightOff, size_t(f->numWeights) * sizeof(float), weights, size_t(nc) * sizeof(float),
                             size_t(nc) * sizeof(float), f->B, cudaMemcpyHostToDevice, f->stream));
  MB2_CUDA(cudaStreamSynchronize(f->stream));
  return MB2_OK;
}

// per-instance weights of one block from DEVICE memory [B][nc] (the torch binding's path: no host round trip)
int mb2_set_constraint_weights_device(mb2_solver_function* f, int32_t index, const float* weights_device, void* stream) {
  MB2_CHECK(f != nullptr && index >= 0 && index < int(f->efs.size()) && weights_device, "invalid constraint weights");
  HostErrorFunction& ef = f->efs[index];
  MB2_CHECK(ef.kind <= 2 || ef.kind == 5, "constraint weights apply to Position/Orientation/Plane error functions");
  MB2_DEVICE_GUARD(f->ch->device);
  const int nc = ef.numConstraints();
  if (!f->weightsPerInstance) { // expand the shared array to [B][numWeights]
    std::vector<float> all(size_t(f->B) * f->numWeights);
    for (int b = 0; b < f->B; ++b) std::copy(f->hWeights.begin(), f->hWeights.end(), all.begin() + size_t(b) * f->numWeights);
    f->dWeights.release();
    MB2_CUDA(f->dWeights.upload(all, f->stream));
    MB2_CUDA(cudaStreamSynchronize(f->stream));
    f->weightsPerInstance = true;
    f->planDirty = true; // tables() carries the per-instance flag
  }
  cudaStream_t st = stream ? (cudaStream_t)stream : f->stream;
  if (nc > 0)
    MB2_CUDA(cudaMemcpy2DAsync(f->dWeights.p + ef.weightOff, size_t(f->numWeights) * sizeof(float), weights_device, size_t(nc) * sizeof(float), size_t(nc) * sizeof(float), f->B,
                               cudaMemcpyDeviceToDevice, st));
  return MB2_OK;
}

int mb2_solver_function_set_enabled_parameters(mb2_solver_function* f, const uint64_t* bits) {
  MB2_CHECK(f != nullptr && bits != nullptr, "null argument");
  bitsToEnabled(bits, f->ch->host.numParams, f->enabled);
  f->planDirty = true;
  return MB2_OK;
}

int mb2_solver_function_get_error(mb2_solver_function* f, const float* params, double* errors) {
  MB2_CHECK(f != nullptr && params && errors, "null argument");
  MB2_DEVICE_GUARD(f->ch->device);
  int rc = ensurePlan(f, f->planMode, f->planSchedDense, f->planAlignRows);
  if (rc != MB2_OK) return rc;
  const size_t n = f->ch->host.numParams;
  MB2_CUDA(cudaMemcpyAsync(f->dTheta.p, params, size_t(f->B) * n * sizeof(float), cudaMemcpyHostToDevice, f->stream));
  MB2_CUDA(launchSweep(sweepArgs(f, f->dTheta.p, nullptr), false, f->stream));
  MB2_CUDA(cudaMemcpyAsync(errors, f->dErrors.p, size_t(f->B) * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  MB2_CUDA(cudaStreamSynchronize(f->stream));
  return MB2_OK;
}

int mb2_solver_function_get_jacobian(mb2_solver_function* f, const float* params, float* jac, float* residual, double* errors, int32_t* actualRows) {
  MB2_CHECK(f != nullptr && params, "null argument");
  MB2_DEVICE_GUARD(f->ch->device);
  int rc = ensurePlan(f, 0);
  if (rc != MB2_OK) return rc;
  const size_t n = f->ch->host.numParams;
  const int rows = mb2_solver_function_jacobian_rows(f);
  MB2_CUDA(cudaMemcpyAsync(f->dTheta.p, params, size_t(f->B) * n * sizeof(float), cudaMemcpyHostToDevice, f->stream));
  MB2_CUDA(launchSweep(sweepArgs(f, f->dTheta.p, nullptr), true, f->stream));
  for (int b = 0; b < f->B && rows > 0; ++b) { // parity/debug entry point: one strided copy per instance
    const float* Jb = f->dJ.p + size_t(b) * (n + 1) * f->ldJ;
    if (jac)
      MB2_CUDA(cudaMemcpy2DAsync(jac + size_t(b) * n * rows, size_t(rows) * sizeof(float), Jb, size_t(f->ldJ) * sizeof(float), size_t(rows) * sizeof(float), n,
                                 cudaMemcpyDeviceToHost, f->stream));
    if (residual)
      MB2_CUDA(cudaMemcpyAsync(residual + size_t(b) * rows, Jb + n * f->ldJ, size_t(rows) * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
  }
  if (errors) MB2_CUDA(cudaMemcpyAsync(errors, f->dErrors.p, size_t(f->B) * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  MB2_CUDA(cudaStreamSynchronize(f->stream));
  if (actualRows) *actualRows = rows; // solver_function.cpp:50 actualRows = totalRows (padded)
  return MB2_OK;
}

// getJacobian with parameters and outputs resident on the device (backward pass of the torch binding): jacobian [B][n][ldJ] in the
// device layout (column c of instance b at (b * (n + 1) + c) * ldJ, ldJ = mb2_solver_function_jacobian_stride), residual = column n.
int mb2_solver_function_get_jacobian_device(mb2_solver_function* f, const float* params_device, const float** jacobian_device, int32_t* ld, void* stream) {
  MB2_CHECK(f != nullptr && params_device && jacobian_device, "null argument");
  MB2_DEVICE_GUARD(f->ch->device);
  int rc = ensurePlan(f, 0);
  if (rc != MB2_OK) return rc;
  cudaStream_t st = stream ? (cudaStream_t)stream : f->stream;
  MB2_CUDA(launchSweep(sweepArgs(f, params_device, nullptr), true, st));
  *jacobian_device = f->dJ.p;
  if (ld) *ld = f->ldJ;
  return MB2_OK;
}

int mb2_solver_function_get_jtjr(mb2_solver_function* f, const float* params, int32_t jtjMode, float* jtj, float* jtr, double* errors) {
  MB2_CHECK(f != nullptr && params, "null argument");
  MB2_DEVICE_GUARD(f->ch->device);
  int rc = ensurePlan(f, 0);
  if (rc != MB2_OK) return rc;
  const size_t n = f->ch->host.numParams;
  const int ap = f->plan.actualParameters;
  MB2_CHECK(ap > 0, "no enabled parameters");
  const int mode = resolveJtjMode(f, jtjMode, ap);
  if (mode < 0) return fail(MB2_ERR_UNSUPPORTED, "tensor-core JtJ does not support this shape");
  const int ldH = roundUp(ap + 1, 16);
  const size_t hElems = size_t(f->B) * (ap + 1) * ldH;
  MB2_CUDA(f->dH.resize(hElems));
  MB2_CUDA(cudaMemcpyAsync(f->dTheta.p, params, size_t(f->B) * n * sizeof(float), cudaMemcpyHostToDevice, f->stream));
  MB2_CUDA(launchSweep(sweepArgs(f, f->dTheta.p, nullptr), true, f->stream));
  MB2_CUDA(cudaMemsetAsync(f->dH.p, 0, hElems * sizeof(float), f->stream));
  rc = runJtJ(f, mode, ap, f->dH.p, ldH, size_t(ap + 1) * ldH, nullptr, f->stream);
  if (rc != MB2_OK) return rc;
  std::vector<float> h(hElems);
  MB2_CUDA(cudaMemcpyAsync(h.data(), f->dH.p, hElems * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
  if (errors) MB2_CUDA(cudaMemcpyAsync(errors, f->dErrors.p, size_t(f->B) * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  MB2_CUDA(cudaStreamSynchronize(f->stream));
  for (int b = 0; b < f->B; ++b) { // device layout is the full symmetric [JtJ, Jtr]; the ABI returns the lower triangle + Jtr
    const float* Hb = h.data() + size_t(b) * (ap + 1) * ldH;
    for (int j = 0; j < ap; ++j) {
      if (jtj) for (int i = j; i < ap; ++i) jtj[(size_t(b) * ap + i) * ap + j] = Hb[size_t(j) * ldH + i];
      if (jtr) jtr[size_t(b) * ap + j] = Hb[size_t(j) * ldH + ap];
    }
  }
  return MB2_OK;
}

int mb2_solver_function_get_skeleton_state(mb2_solver_function* f, const float* params, float* state) {
  MB2_CHECK(f != nullptr && params && state, "null argument");
  MB2_DEVICE_GUARD(f->ch->device);
  int rc = ensurePlan(f, f->planMode, f->planSchedDense, f->planAlignRows);
  if (rc != MB2_OK) return rc;
  const size_t n = f->ch->host.numParams;
  const size_t sz = size_t(f->B) * f->ch->host.numJoints * 8;
  MB2_CUDA(f->dState.resize(sz));
  MB2_CUDA(cudaMemcpyAsync(f->dTheta.p, params, size_t(f->B) * n * sizeof(float), cudaMemcpyHostToDevice, f->stream));
  SweepArgs a = sweepArgs(f, f->dTheta.p, nullptr);
  a.stateOut = f->dState.p;
  MB2_CUDA(launchSweep(a, false, f->stream));
  MB2_CUDA(cudaMemcpyAsync(state, f->dState.p, sz * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
  MB2_CUDA(cudaStreamSynchronize(f->stream));
  return MB2_OK;
}

// ---------------------------------------------------------------------------------------------
// Solver
// ---------------------------------------------------------------------------------------------
int mb2_solver_create(mb2_solver_function* f, const mb2_gauss_newton_options* opt, mb2_solver** out) {
  MB2_CHECK(f != nullptr && out != nullptr, "null argument");
  auto s = std::make_unique<mb2_solver>();
  s->fn = f;
  if (opt) s->opt = *opt; else mb2_default_gauss_newton_options(&s->opt);
  MB2_DEVICE_GUARD(f->ch->device);
  MB2_CUDA(cudaMallocHost(&s->hActiveCount, sizeof(int)));
  *out = s.release();
  return MB2_OK;
}
void mb2_solver_destroy(mb2_solver* s) { delete s; }
int mb2_solver_set_options(mb2_solver* s, const mb2_gauss_newton_options* opt) {
  MB2_CHECK(s != nullptr && opt != nullptr, "null argument");
  s->opt = *opt;
  return MB2_OK;
}
int mb2_solver_set_enabled_parameters(mb2_solver* s, const uint64_t* bits) {
  MB2_CHECK(s != nullptr, "null solver");
  return mb2_solver_function_set_enabled_parameters(s->fn, bits);
}
int mb2_solver_set_profiling(mb2_solver* s, int32_t enabled) {
  MB2_CHECK(s != nullptr, "null solver");
  s->profiling = enabled != 0;
  s->inKernelProfile = enabled >= 2;
  return MB2_OK;
}

int mb2_solver_solve_device(mb2_solver* s, float* theta, void* cudaStream) {
  MB2_CHECK(s != nullptr && theta != nullptr, "null argument");
  mb2_solver_function* f = s->fn;
  MB2_DEVICE_GUARD(f->ch->device);
  NvtxRange nvtxSolve("mb2::SolverT::solve");
  const mb2_gauss_newton_options& o = s->opt;
  // Cholesky path: 0 auto, 1 dense Eigen-structured kernel, 2 tile schedule on the dense pattern, 3 tile schedule on the sparse pattern
  int numEnabled = 0;
  for (uint8_t e : f->enabled) numEnabled += e ? 1 : 0;
  int cholMode = o.cholesky_mode;
  if (cholMode == 0) cholMode = numEnabled >= 48 ? 3 : 1;
  // GaussNewtonSolverQRT's step: Householder QR of the K-major Jacobian (ik_qr.cuh); TrustRegionQRT's iteration builds on the same fold (ik_tr_qr.cuh)
  const bool useTr = o.linear_solver == MB2_LINEAR_SOLVER_TRUST_REGION_QR;
  const bool useQr = o.linear_solver == MB2_LINEAR_SOLVER_QR || useTr;
  if (useTr && !(o.trust_region_radius > 0.f)) return fail(MB2_ERR_INVALID_ARGUMENT, "trust_region_radius must be positive");
  if (useQr) {
    if (o.jtj_mode == MB2_JTJ_SPARSE_TILES || o.cholesky_mode >= 2 || o.fused_mode >= MB2_FUSED_PERSISTENT)
      return fail(MB2_ERR_UNSUPPORTED, "the QR step works on the dense Jacobian: it excludes the tile-sparse Gram, the tile Cholesky and the fused kernels");
    cholMode = 1;
  }
  // tile-sparse Gram: default whenever the tile-scheduled Cholesky runs (MB2_JTJ_AUTO), or on request
  bool useGram = cholMode >= 2 && (o.jtj_mode == MB2_JTJ_AUTO || o.jtj_mode == MB2_JTJ_SPARSE_TILES);
  if (o.jtj_mode == MB2_JTJ_SPARSE_TILES && cholMode < 2) return fail(MB2_ERR_UNSUPPORTED, "MB2_JTJ_SPARSE_TILES needs the tile-scheduled Cholesky");
  int rc = ensurePlan(f, cholMode >= 2 ? 2 : 1, cholMode == 2, useGram);
  if (rc != MB2_OK) return rc;
  bool useSchedule = cholMode >= 2;
  if (useSchedule && choleskyScheduledSmemBytes(f->plan.numCols, f->sched->host.nPad, f->sched->host.numTiles, f->sched->dev.blobInts) > size_t(200 * 1024)) {
    if (o.cholesky_mode >= 2) return fail(MB2_ERR_UNSUPPORTED, "tile schedule does not fit in shared memory for this system");
    useSchedule = false; // fall back to the dense kernel (matrix in global memory)
    useGram = false;
    rc = ensurePlan(f, 1);
    if (rc != MB2_OK) return rc;
  }
  cudaStream_t st = cudaStream ? (cudaStream_t)cudaStream : f->stream;
  const int B = f->B, n = f->ch->host.numParams;
  const int ns = f->plan.numCols;
  const int maxIt = int(std::min<uint64_t>(o.max_iterations, 1u << 30));
  const int minIt = int(std::min<uint64_t>(o.min_iterations, 1u << 30));
  MB2_CHECK(ns > 0, "no enabled parameters");
  if (useGram && gramTilesSmemBytes(size_t(f->sched->gram.stride), f->sched->gBlobInts) > size_t(200 * 1024)) {
    if (o.jtj_mode == MB2_JTJ_SPARSE_TILES) return fail(MB2_ERR_UNSUPPORTED, "Jacobian strips do not fit in shared memory for this system");
    useGram = false;
    rc = ensurePlan(f, 2, cholMode == 2, false);
    if (rc != MB2_OK) return rc;
  }
  // ---- fused persistent kernel: the whole solve in one launch (ik_fused.cuh) ----
  s->lastFused = false;
  {
    const bool eligible = useGram && useSchedule && o.do_line_search == 0 && f->sched->fConfig.groups >= 1 && f->sched->fBlob.p != nullptr;
    if (o.fused_mode == MB2_FUSED_PERSISTENT && !eligible)
      return fail(MB2_ERR_UNSUPPORTED, "the fused kernel needs the tile-sparse Gram + tile-scheduled Cholesky path, no line search, and a plan that fits in shared memory");
    // AUTO takes the persistent kernel when the whole batch is one wave of instance groups: a single launch with no host round trip
    // beats ~2 launches per iteration there (measured 0.60 vs 0.89 ms at 256 x humanoid72 x 10 iterations); at thousands of instances
    // the per-iteration kernels hide latency better (three instances per SM cannot) and AUTO keeps them
    int smCount = 0;
    if (eligible && o.fused_mode == MB2_FUSED_AUTO) {
      int dev0 = 0;
      MB2_CUDA(cudaGetDevice(&dev0));
      MB2_CUDA(cudaDeviceGetAttribute(&smCount, cudaDevAttrMultiProcessorCount, dev0));
    }
    const bool oneWave = eligible && o.fused_mode == MB2_FUSED_AUTO && B <= f->sched->fConfig.groups * smCount;
    if ((o.fused_mode == MB2_FUSED_PERSISTENT || oneWave) && eligible) {
      const DeviceSchedule& ds = *f->sched;
      MB2_CUDA(s->dIterations.resize(B));
      MB2_CUDA(s->dStatus.resize(B));
      MB2_CUDA(s->dWorkCounter.resize(1));
      MB2_CUDA(cudaMemsetAsync(s->dWorkCounter.p, 0, sizeof(int32_t), st));
      if (o.store_error_history) {
        MB2_CUDA(s->dHistory.resize(size_t(B) * std::max(maxIt, 1)));
        MB2_CUDA(cudaMemsetAsync(s->dHistory.p, 0, size_t(B) * std::max(maxIt, 1) * sizeof(double), st));
        s->historyStride = size_t(std::max(maxIt, 1));
      }
      FusedArgs fa{};
      fa.batch = B;
      fa.T = f->tables();
      fa.blob = ds.fBlob.p;
      fa.L = ds.fLayout;
      fa.S = ds.dev;
      fa.schedGlobal = ds.dev.blob;
      for (int k = 0; k < 8; ++k) fa.gramOff[k] = ds.gOffsets[k];
      fa.numStrips = ds.gram.numStrips;
      fa.stripStride = ds.gram.stride;
      fa.residOff = ds.gram.residOff;
      fa.numOrder = int32_t(ds.gram.tileOrder.size());
      fa.regularization = o.regularization;
      fa.threshold = o.threshold;
      fa.minIterations = minIt;
      fa.maxIterations = maxIt;
      fa.theta = theta;
      fa.ldTheta = n;
      fa.targets = f->dTargets.p;
      fa.cweights = f->dWeights.p;
      fa.errors = f->dErrors.p;
      fa.iterations = s->dIterations.p;
      fa.status = s->dStatus.p;
      fa.history = o.store_error_history ? s->dHistory.p : nullptr;
      fa.workCounter = s->dWorkCounter.p;
      fa.phaseCycles = nullptr;
      if (s->profiling) {
        MB2_CUDA(s->dPhaseCycles.resize(16));
        MB2_CUDA(cudaMemsetAsync(s->dPhaseCycles.p, 0, 16 * sizeof(unsigned long long), st));
        fa.phaseCycles = s->inKernelProfile ? s->dPhaseCycles.p : nullptr;
        if (!s->fusedStart) { MB2_CUDA(cudaEventCreate(&s->fusedStart)); MB2_CUDA(cudaEventCreate(&s->fusedStop)); }
        MB2_CUDA(cudaEventRecord(s->fusedStart, st));
      }
      int dev = 0, sms = 0;
      MB2_CUDA(cudaGetDevice(&dev));
      MB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      {
        NvtxRange nvtxIt("mb2::fusedSolveKernel (GaussNewtonSolverT::doIteration x maxIterations)");
        MB2_CUDA(launchFusedSolve(fa, ds.fConfig, sms, s->inKernelProfile, st));
      }
      if (s->profiling) MB2_CUDA(cudaEventRecord(s->fusedStop, st));
      for (auto& e : s->events) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
      s->events.clear();
      s->kernelLaunches = 1;
      s->lastFused = true;
      s->lastFusedGroups = ds.fConfig.groups;
      return MB2_OK;
    }
  }
  // ---- Gram + Cholesky in one launch per iteration (gramCholeskyKernel): the default on the tile path ----
  bool useGramChol = false;
  if (useGram && useSchedule) {
    const DeviceSchedule& ds = *f->sched;
    const int rounds = std::max(int(ds.gram.tileOrder.size()) / kGramWarps, 1);
    const bool fits = gramCholeskySmemBytes(size_t(ds.gram.stride), ds.gBlobInts, ns, ds.host.nPad, ds.host.numTiles, ds.dev.blobInts) <= size_t(200 * 1024) && 16 * rounds <= 128;
    if (o.fused_mode == MB2_FUSED_GRAM_CHOLESKY && !fits) return fail(MB2_ERR_UNSUPPORTED, "Gram + Cholesky fusion: strips / tiles of this plan do not fit in shared memory or TMEM");
    useGramChol = fits && (o.fused_mode == MB2_FUSED_AUTO || o.fused_mode == MB2_FUSED_GRAM_CHOLESKY);
  } else if (o.fused_mode == MB2_FUSED_GRAM_CHOLESKY) {
    return fail(MB2_ERR_UNSUPPORTED, "Gram + Cholesky fusion needs the tile-sparse Gram + tile-scheduled Cholesky path");
  }
  s->lastGramChol = useGramChol;
  if (useGramChol && s->profiling) {
    MB2_CUDA(s->dPhaseCycles.resize(16));
    MB2_CUDA(cudaMemsetAsync(s->dPhaseCycles.p, 0, 16 * sizeof(unsigned long long), st));
  }
  const int mode = useGram ? MB2_JTJ_SPARSE_TILES : resolveJtjMode(f, o.jtj_mode == MB2_JTJ_SPARSE_TILES ? MB2_JTJ_AUTO : o.jtj_mode, ns);
  if (mode < 0) return fail(MB2_ERR_UNSUPPORTED, "tensor-core JtJ does not support this shape");
  // normal equations: full symmetric [ns+1][ldH] in device-column order (row/column ns = J^T r)
  const int ldH = roundUp(ns + 1, 16);
  const size_t hStride = size_t(ns + 1) * ldH;
  if (!useGram && !useQr) MB2_CUDA(s->dH.resize(size_t(B) * hStride));
  int qrMaxRows = 0, qrChunks = 0;
  if (useQr) { // row chunks: one per error-function block (the reference adds block by block), split when a block does not fit beside R
    qrMaxRows = qrMaxChunkRows(ns, size_t(200 * 1024));
    if (qrMaxRows < 8) return fail(MB2_ERR_UNSUPPORTED, "the QR step keeps R in shared memory: too many enabled parameters for this kernel");
    if (useTr) { // R and the damping rows D (both packed triangles) + the in-kernel getError scratch; the Jacobian chunk shares D's storage
      qrMaxRows = std::min(qrMaxRows, std::max(8, ns / 2));
      if (trQrSmemFloats(ns, n, f->ch->host.numJoints, qrMaxRows) * sizeof(float) > size_t(220 * 1024))
        return fail(MB2_ERR_UNSUPPORTED, "the trust-region QR kernel keeps R and the damping rows in shared memory: too many enabled parameters / joints");
    }
    std::vector<int32_t> starts{0};
    int r0 = 0, widest = 0;
    for (const auto& ef : f->efs) {
      if (!(ef.weight > 0.f)) continue;
      const int size = jacobianBlockSize(f->ch->host, ef);
      for (int done = 0; done < size; done += qrMaxRows) { const int p = std::min(qrMaxRows, size - done); starts.push_back(r0 + done + p); widest = std::max(widest, p); }
      r0 += size;
    }
    qrChunks = int(starts.size()) - 1;
    qrMaxRows = std::max(widest, 1);
    MB2_CUDA(s->dQrChunks.upload(starts, st));
    if (useTr) { // TrustRegionQRT::initializeSolver (trust_region_qr.cpp:38-40): the current radius starts at the option's value
      MB2_CUDA(s->dRadius.upload(std::vector<float>(size_t(B), o.trust_region_radius), st));
      MB2_CUDA(s->dRSaved.resize(size_t(B) * ((size_t(ns) * (ns + 1) / 2 + 3) & ~size_t(3))));
    }
  }
  float* Hbuf = s->dH.p;
  const size_t tilesStride = useGram ? size_t(f->sched->host.numTiles) * 256 + f->sched->host.nPad : 0;
  if (useGram && !(o.fused_mode == MB2_FUSED_AUTO || o.fused_mode == MB2_FUSED_GRAM_CHOLESKY)) MB2_CUDA(s->dTiles.resize(size_t(B) * tilesStride));
  const int ldG = cholGradientLd(ns);
  MB2_CUDA(s->dGrad.resize(size_t(B) * ldG));
  MB2_CUDA(s->dDelta.resize(size_t(B) * ns));
  MB2_CUDA(s->dTheta0.resize(size_t(B) * n));
  MB2_CUDA(s->dLastErrors.resize(B));
  MB2_CUDA(s->dActive.resize(B));
  MB2_CUDA(s->dIterations.resize(B));
  MB2_CUDA(s->dStatus.resize(B));
  MB2_CUDA(s->dActiveCount.resize(1));
  const bool lineSearch = o.do_line_search != 0 && !useTr; // (TrustRegionQRT has no line search: steps are accepted or rejected by rho)
  if (lineSearch) {
    MB2_CUDA(s->dThetaOrig.resize(size_t(B) * n));
    MB2_CUDA(s->dTrialErrors.resize(B));
    MB2_CUDA(s->dScale.resize(B));
    MB2_CUDA(s->dSearching.resize(B));
    MB2_CUDA(s->dGradDotDelta.resize(B));
  }
  if (o.store_error_history) {
    MB2_CUDA(s->dHistory.resize(size_t(B) * std::max(maxIt, 1)));
    MB2_CUDA(cudaMemsetAsync(s->dHistory.p, 0, size_t(B) * std::max(maxIt, 1) * sizeof(double), st));
    s->historyStride = size_t(std::max(maxIt, 1));
  }
  for (auto& e : s->events) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
  s->events.clear();
  s->kernelLaunches = 0;

  initSolveStateKernel<<<(B + 127) / 128, 128, 0, st>>>(B, s->dActive.p, s->dIterations.p, s->dStatus.p, s->dLastErrors.p, f->dErrors.p);
  MB2_CUDA(cudaGetLastError());
  MB2_CUDA(cudaMemcpyAsync(s->dTheta0.p, theta, size_t(B) * n * sizeof(float), cudaMemcpyDeviceToDevice, st));

  static const bool cholProfile = getenv("MB2_CHOL_PROFILE") != nullptr;
  constexpr int kPollEvery = 4; // iterations between two reads of the device-side active counter (each read blocks the host on `st`)
  for (int it = 0; it < maxIt; ++it) {
    MB2_CUDA(cudaMemsetAsync(s->dActiveCount.p, 0, sizeof(int), st));
    // --- doIteration (gauss_newton_solver.cpp:224-280) ---
    recordPhaseStart(s, 0, st);
    MB2_CUDA(launchSweep(sweepArgs(f, theta, s->dActive.p), true, st));
    recordPhaseStop(s, st);
    GramArgs g{};
    if (!useGramChol && !useQr) recordPhaseStart(s, 1, st);
    if (useQr) {
      // nothing: the QR kernel reads the Jacobian directly
    } else if (useGram) {
      const DeviceSchedule& ds = *f->sched;
      g.batch = B;
      g.strips = f->dJ.p;
      g.stripStride = size_t(ds.gram.stride);
      g.residOff = ds.gram.residOff;
      g.active = s->dActive.p;
      g.numStrips = ds.gram.numStrips;
      g.numTiles = ds.host.numTiles;
      g.numTileCols = ds.host.numTileCols;
      g.nPad = ds.host.nPad;
      g.numOrder = int32_t(ds.gram.tileOrder.size());
      g.blob = ds.gBlob.p;
      g.blobInts = ds.gBlobInts;
      g.offTileOrder = ds.gOffsets[0]; g.offTilePairStart = ds.gOffsets[1]; g.offPairA = ds.gOffsets[2]; g.offPairB = ds.gOffsets[3];
      g.offColStripStart = ds.gOffsets[4]; g.offColStrip = ds.gOffsets[5]; g.offStripRow = ds.gOffsets[6]; g.offTileInfo = ds.gOffsets[7];
      g.regularization = o.regularization;
      g.out = s->dTiles.p;
      g.outStride = tilesStride;
      if (!useGramChol) {
        if (s->dTiles.n < size_t(B) * tilesStride) MB2_CUDA(s->dTiles.resize(size_t(B) * tilesStride));
        g.out = s->dTiles.p;
        MB2_CUDA(launchGramTiles(g, st));
      }
    } else {
      rc = runJtJ(f, mode, ns, Hbuf, ldH, hStride, s->dActive.p, st, s->dGrad.p, ldG);
      if (rc != MB2_OK) return rc;
    }
    if (!useGramChol && !useQr) recordPhaseStop(s, st);
    CholArgs c{};
    c.batch = B;
    c.H = Hbuf;
    c.ns = ns;
    c.ldH = ldH;
    c.hStride = hStride;
    c.regularization = o.regularization;
    c.cols = f->dDeviceCols.p;
    c.theta = theta;
    c.ldTheta = n;
    c.delta = s->dDelta.p;
    c.applyUpdate = lineSearch ? 0 : 1;
    c.errors = f->dErrors.p;
    c.lastErrors = s->dLastErrors.p;
    c.active = s->dActive.p;
    c.iterations = s->dIterations.p;
    c.status = s->dStatus.p;
    c.history = o.store_error_history ? s->dHistory.p : nullptr;
    c.iteration = it;
    c.minIterations = minIt;
    c.maxIterations = maxIt;
    c.threshold = o.threshold;
    c.activeCount = s->dActiveCount.p;
    c.bookkeeping = lineSearch ? 0 : 1;
    c.gradDotDelta = lineSearch ? s->dGradDotDelta.p : nullptr;
    c.g = s->dGrad.p;
    c.ldG = ldG;
    c.tilesIn = useGram ? s->dTiles.p : nullptr;
    c.tilesStride = tilesStride;
    c.profile = (it == 0 && cholProfile) ? 1 : 0;
    recordPhaseStart(s, 2, st);
    if (useQr) {
      QrArgs q{};
      q.c = c;
      q.jacobian = f->dJ.p;
      q.numCols = f->plan.numCols;
      q.ldJ = f->ldJ;
      q.chunkStart = s->dQrChunks.p;
      q.numChunks = qrChunks;
      if (useTr) {
        TrQrArgs t{};
        t.q = q;
        t.T = f->tables();
        t.targets = f->dTargets.p;
        t.cweights = f->dWeights.p;
        t.radius = s->dRadius.p;
        t.rSaved = s->dRSaved.p;
        t.maxRadius = 10.f; // trust_region_qr.h:73
        t.maxChunkRows = qrMaxRows;
        MB2_CUDA(launchTrustRegionQr(t, st));
      } else {
        MB2_CUDA(launchQrSolve(q, qrMaxRows, st));
      }
    } else if (useGramChol) {
      GramCholArgs gc{};
      gc.g = g;
      gc.c = c;
      gc.c.tilesIn = nullptr;
      gc.phaseCycles = s->inKernelProfile ? s->dPhaseCycles.p : nullptr;
      MB2_CUDA(launchGramCholesky(gc, f->sched->dev, s->inKernelProfile, st));
    } else if (useSchedule) MB2_CUDA(launchCholeskyScheduled(c, f->sched->dev, st));
    else MB2_CUDA(launchCholesky(c, st));
    recordPhaseStop(s, st);
    if (lineSearch) { // gauss_newton_solver.cpp:283-313 / subset_gauss_newton_solver.cpp:119-141
      MB2_CUDA(cudaMemcpyAsync(s->dThetaOrig.p, theta, size_t(B) * n * sizeof(float), cudaMemcpyDeviceToDevice, st));
      initLineSearchKernel<<<(B + 127) / 128, 128, 0, st>>>(B, s->dActive.p, s->dSearching.p, s->dScale.p);
      MB2_CUDA(cudaGetLastError());
      for (int step = 0; step < 10; ++step) {
        MB2_CUDA(launchTrialUpdate(B, s->dThetaOrig.p, n, s->dDelta.p, ns, f->dDeviceCols.p, s->dScale.p, theta, s->dSearching.p, st));
        SweepArgs ea = sweepArgs(f, theta, s->dSearching.p);
        ea.errors = s->dTrialErrors.p;
        recordPhaseStart(s, 3, st);
        MB2_CUDA(launchSweep(ea, false, st));
        recordPhaseStop(s, st);
        LineSearchArgs la{};
        la.batch = B; la.ns = ns; la.numParams = n; la.ldTheta = n;
        la.errors = f->dErrors.p;
        la.trialErrors = s->dTrialErrors.p;
        la.gradDotDelta = s->dGradDotDelta.p;
        la.scale = s->dScale.p;
        la.searching = s->dSearching.p;
        la.step = step;
        la.subsetVariant = o.subset_line_search;
        la.active = s->dActive.p;
        MB2_CUDA(launchLineSearchStep(la, st));
        s->kernelLaunches += 2;
      }
      BookkeepingArgs ba{};
      ba.batch = B;
      ba.errors = f->dErrors.p; ba.lastErrors = s->dLastErrors.p; ba.active = s->dActive.p; ba.iterations = s->dIterations.p;
      ba.status = s->dStatus.p; ba.history = o.store_error_history ? s->dHistory.p : nullptr;
      ba.theta = theta; ba.ldTheta = n; ba.numParams = n;
      ba.iteration = it; ba.minIterations = minIt; ba.maxIterations = maxIt; ba.threshold = o.threshold; ba.activeCount = s->dActiveCount.p;
      MB2_CUDA(launchBookkeeping(ba, st));
      s->kernelLaunches += 1;
    }
    // Instances can only stop once iteration >= minIterations (solver.cpp:113): poll the device counter from then on (every
    // kPollEvery-th iteration: converged instances cost nothing on the device, the poll costs a host round trip) so that a
    // converged batch does not run to maxIterations.
    if (it + 1 < maxIt && it >= minIt && (it - minIt) % kPollEvery == kPollEvery - 1) {
      MB2_CUDA(cudaMemcpyAsync(s->hActiveCount, s->dActiveCount.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      MB2_CUDA(cudaStreamSynchronize(st));
      if (*s->hActiveCount == 0) break;
    }
  }
  finalizeKernel<<<B, 128, 0, st>>>(B, n, theta, s->dTheta0.p, s->dStatus.p);
  MB2_CUDA(cudaGetLastError());
  s->kernelLaunches += 2;
  return MB2_OK;
}

static void collectProfile(mb2_solver* s) {
  if (s->profiling && s->lastFused && s->fusedStart) {
    float ms = 0.f;
    s->fusedMs = cudaEventElapsedTime(&ms, s->fusedStart, s->fusedStop) == cudaSuccess ? double(ms) : 0.0;
  }
  if (s->profiling) {
    for (int k = 0; k < 4; ++k) { s->phaseMs[k] = 0; s->phaseLaunches[k] = 0; }
    for (auto& e : s->events) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e.start, e.stop) == cudaSuccess) { s->phaseMs[e.phase] += ms; s->phaseLaunches[e.phase]++; }
    }
  }
}

int mb2_solver_get_results(mb2_solver* s, double* errors, int32_t* iterations, int32_t* status) {
  MB2_CHECK(s != nullptr, "null solver");
  mb2_solver_function* f = s->fn;
  const int B = f->B;
  MB2_DEVICE_GUARD(f->ch->device);
  MB2_CUDA(cudaDeviceSynchronize());
  if (errors) MB2_CUDA(cudaMemcpy(errors, f->dErrors.p, size_t(B) * sizeof(double), cudaMemcpyDeviceToHost));
  std::vector<int32_t> its(B);
  MB2_CUDA(cudaMemcpy(its.data(), s->dIterations.p, size_t(B) * sizeof(int32_t), cudaMemcpyDeviceToHost));
  s->totalIterations = 0;
  for (int v : its) s->totalIterations += uint64_t(v);
  if (iterations) std::copy(its.begin(), its.end(), iterations);
  if (status) MB2_CUDA(cudaMemcpy(status, s->dStatus.p, size_t(B) * sizeof(int32_t), cudaMemcpyDeviceToHost));
  collectProfile(s);
  return MB2_OK;
}

int mb2_solver_solve_async(mb2_solver* s, float* params) {
  MB2_CHECK(s != nullptr && params != nullptr, "null argument");
  mb2_solver_function* f = s->fn;
  MB2_DEVICE_GUARD(f->ch->device);
  const size_t B = size_t(f->B);
  const size_t bytes = B * f->ch->host.numParams * sizeof(float);
  MB2_CUDA(s->dThetaStage.resize(B * f->ch->host.numParams));
  MB2_CUDA(cudaMemcpyAsync(s->dThetaStage.p, params, bytes, cudaMemcpyHostToDevice, f->stream));
  s->resultsStaged = false;
  int rc = mb2_solver_solve_device(s, s->dThetaStage.p, f->stream);
  if (rc != MB2_OK) return rc;
  MB2_CUDA(cudaMemcpyAsync(params, s->dThetaStage.p, bytes, cudaMemcpyDeviceToHost, f->stream));
  // the per-instance results follow the parameters on the same stream into pinned staging (one synchronisation in mb2_solver_wait)
  const size_t need = B * (sizeof(double) + 2 * sizeof(int32_t));
  if (s->hResultsBytes < need) {
    if (s->hResults) cudaFreeHost(s->hResults);
    s->hResults = nullptr;
    s->hResultsBytes = 0;
    MB2_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->hResults), need));
    s->hResultsBytes = need;
  }
  MB2_CUDA(cudaMemcpyAsync(s->hResults, f->dErrors.p, B * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  MB2_CUDA(cudaMemcpyAsync(s->hResults + B * sizeof(double), s->dIterations.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, f->stream));
  MB2_CUDA(cudaMemcpyAsync(s->hResults + B * (sizeof(double) + sizeof(int32_t)), s->dStatus.p, B * sizeof(int32_t), cudaMemcpyDeviceToHost, f->stream));
  s->resultsStaged = true;
  return MB2_OK;
}

int mb2_solver_wait(mb2_solver* s, double* errors, int32_t* iterations, int32_t* status) {
  MB2_CHECK(s != nullptr, "null solver");
  MB2_DEVICE_GUARD(s->fn->ch->device);
  MB2_CUDA(cudaStreamSynchronize(s->fn->stream));
  if (!s->resultsStaged) return mb2_solver_get_results(s, errors, iterations, status);
  const size_t B = size_t(s->fn->B);
  const int32_t* its = reinterpret_cast<const int32_t*>(s->hResults + B * sizeof(double));
  if (errors) std::memcpy(errors, s->hResults, B * sizeof(double));
  s->totalIterations = 0;
  for (size_t b = 0; b < B; ++b) s->totalIterations += uint64_t(its[b]);
  if (iterations) std::memcpy(iterations, its, B * sizeof(int32_t));
  if (status) std::memcpy(status, s->hResults + B * (sizeof(double) + sizeof(int32_t)), B * sizeof(int32_t));
  s->resultsStaged = false;
  collectProfile(s);
  return MB2_OK;
}

int mb2_solver_solve(mb2_solver* s, float* params, double* errors, int32_t* iterations, int32_t* status) {
  const int rc = mb2_solver_solve_async(s, params);
  if (rc != MB2_OK) return rc;
  return mb2_solver_wait(s, errors, iterations, status);
}

int mb2_solver_get_error_history(mb2_solver* s, double* history) {
  MB2_CHECK(s != nullptr && history != nullptr, "null argument");
  MB2_CHECK(s->opt.store_error_history && s->dHistory.p, "error history was not stored (set store_error_history)");
  MB2_CUDA(cudaMemcpy(history, s->dHistory.p, size_t(s->fn->B) * s->historyStride * sizeof(double), cudaMemcpyDeviceToHost));
  return MB2_OK;
}

int mb2_solver_get_counters(mb2_solver* s, uint64_t* totalIterations, uint64_t* kernelLaunches) {
  MB2_CHECK(s != nullptr, "null solver");
  if (totalIterations) *totalIterations = s->totalIterations;
  if (kernelLaunches) *kernelLaunches = s->kernelLaunches;
  return MB2_OK;
}

int mb2_solver_get_plan_stats(mb2_solver* s, int64_t stats[12]) {
  MB2_CHECK(s != nullptr && stats != nullptr, "null argument");
  const mb2_solver_function* f = s->fn;
  int64_t nnz = 0;
  for (const CellDesc& c : f->plan.cells) nnz += f->plan.units[c.unit].numRows;
  const bool tiles = f->planMode == 2 && f->sched && f->sched->valid;
  stats[0] = nnz;
  stats[1] = f->plan.numCols;
  stats[2] = f->ldJ;
  stats[3] = f->planMode == 0 ? int64_t(f->plan.enabledList.size()) : int64_t(f->plan.numCols);
  stats[4] = tiles ? f->sched->host.numTiles : 0;
  stats[5] = tiles ? f->sched->host.tileOps : 0;
  stats[6] = tiles ? f->sched->host.numLevels : 0;
  stats[7] = f->plan.numRows;
  const bool gram = tiles && f->planAlignRows && f->sched->gramValid;
  stats[8] = gram ? f->sched->gram.stride : 0;
  stats[9] = gram ? f->sched->gram.macs : 0;
  stats[10] = gram ? int64_t(f->sched->gram.pairA.size()) : 0;
  stats[11] = gram ? f->sched->fConfig.groups : 0;
  return MB2_OK;
}

int mb2_solver_get_fused_profile(mb2_solver* s, int32_t* fused, int32_t* groups, double* kernelMs, uint64_t phaseCycles[12]) {
  MB2_CHECK(s != nullptr, "null solver");
  if (fused) *fused = s->lastFused ? 1 : (s->lastGramChol ? 2 : 0);
  if (groups) *groups = s->lastFused ? s->lastFusedGroups : 0;
  if (kernelMs) *kernelMs = s->lastFused ? s->fusedMs : (s->lastGramChol ? s->phaseMs[2] : 0.0);
  if (phaseCycles) {
    for (int k = 0; k < 12; ++k) phaseCycles[k] = 0;
    if ((s->lastFused || s->lastGramChol) && s->profiling && s->dPhaseCycles.p) {
      MB2_DEVICE_GUARD(s->fn->ch->device);
      unsigned long long h[12];
      MB2_CUDA(cudaMemcpy(h, s->dPhaseCycles.p, sizeof(h), cudaMemcpyDeviceToHost));
      if (s->lastFused) {
        for (int k = 0; k < 12; ++k) phaseCycles[k] = h[k];
      } else { // gramCholeskyKernel (block 0, summed over the iterations): prologue, gram, tiles from TMEM, diag, panel, update, backward, finish
        phaseCycles[0] = h[0];
        for (int k = 1; k < 8; ++k) phaseCycles[4 + k] = h[k];
      }
    }
  }
  return MB2_OK;
}

int mb2_solver_get_phase_times(mb2_solver* s, double ms[4], uint64_t launches[4]) {
  MB2_CHECK(s != nullptr, "null solver");
  for (int k = 0; k < 4; ++k) {
    if (ms) ms[k] = s->phaseMs[k];
    if (launches) launches[k] = s->phaseLaunches[k];
  }
  return MB2_OK;
}

} // extern "C"

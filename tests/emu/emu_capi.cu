// TEST HARNESS ONLY — CPU lane-emulation of the device code, built into tests/emu/libmb2_emu.so.
//
// It exports the subset of include/momentum_b200.h that the parity tests use, but runs the
// __host__ __device__ building blocks of momentum_b200/csrc (ik_device.cuh, ik_chol.cuh) and the real
// planner (ik_plan.cpp) lane by lane on the CPU. Purpose: catch planner / indexing / semantics bugs
// in this GPU-less container before spending B200 time. It is not part of the product library and
// nothing in momentum_b200/ loads it.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/momentum_b200.h"
#include "../../momentum_b200/csrc/ik_chol.cuh"
#include "../../momentum_b200/csrc/ik_chol_sched.cuh"
#include "../../momentum_b200/csrc/ik_chol_sched.h"
#include "../../momentum_b200/csrc/ik_device.cuh"
#include "../../momentum_b200/csrc/ik_plan.h"

using namespace mb2;

static thread_local std::string g_err;
static int fail(int c, const std::string& m) { g_err = m; return c; }

struct mb2_character { HostCharacter host; };
struct mb2_solver_function {
  const mb2_character* ch;
  int B;
  std::vector<HostErrorFunction> efs;
  std::vector<uint8_t> enabled;
  int targetStride{0}, numWeights{0};
  bool weightsPerInstance{false};
  std::vector<float> weights, targets;
  Plan plan;
  int ldJ{32};
  std::vector<float> J;
  bool planCompact{false};
  bool stripMode{false};   // Jacobian buffer in strip layout (GramPlan) instead of the K-major matrix
  int residOff{0}, stripStride{0};
  std::vector<double> errors;
};
struct mb2_solver {
  mb2_solver_function* fn;
  mb2_gauss_newton_options opt;
  std::vector<double> errors, history;
  std::vector<int32_t> iterations, status;
  uint64_t totalIterations{0};
};

static int roundUp(int v, int m) { return (v + m - 1) / m * m; }

static std::string plan(mb2_solver_function* f, bool compact) {
  std::string e = buildPlan(f->ch->host, f->efs, f->enabled, compact, f->plan);
  if (!e.empty()) return e;
  f->planCompact = compact;
  f->stripMode = false;
  f->ldJ = std::max(32, roundUp(f->plan.numRows, 32));
  f->J.assign(size_t(f->B) * (f->plan.numCols + 1) * f->ldJ, 0.f);
  f->errors.assign(f->B, 0.0);
  f->targets.resize(size_t(f->B) * std::max(f->targetStride, 1), 0.f);
  if (f->weights.empty()) f->weights.push_back(0.f);
  return "";
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

static FunctionTables tables(const mb2_solver_function* f) {
  FunctionTables T{};
  const HostCharacter& h = f->ch->host;
  T.numJoints = h.numJoints; T.numParams = h.numParams;
  T.parent = h.parent.data(); T.offset = h.offset.data(); T.prerot = h.prerot.data();
  T.ptOuter = h.ptOuter.data(); T.ptInner = h.ptInner.data(); T.ptVals = h.ptVals.data(); T.ptOffsets = h.ptOffsets.data();
  T.numLevels = int(h.levelStart.size()) - 1; T.levelStart = h.levelStart.data(); T.levelJoints = h.levelJoints.data();
  T.numEf = int(f->plan.efs.size()); T.numUnits = int(f->plan.units.size()); T.numCells = int(f->plan.cells.size());
  T.efs = f->plan.efs.data(); T.units = f->plan.units.data(); T.cells = f->plan.cells.data(); T.contribs = f->plan.contribs.data();
  T.limitData = f->plan.limitData.data();
  T.targetStride = f->targetStride; T.recStride = f->plan.recStride; T.numRows = f->plan.numRows; T.ldJ = f->ldJ; T.numCols = f->plan.numCols;
  T.weightsPerInstance = f->weightsPerInstance; T.numWeights = f->numWeights;
  T.stripMode = f->stripMode ? 1 : 0; T.residOff = f->residOff;
  T.jacobianStride = f->stripMode ? size_t(f->stripStride) : size_t(f->plan.numCols + 1) * f->ldJ;
  return T;
}

// emulation of sweepKernel<kJacobian> for one instance (the warp's lanes run in sequence)
template <bool kJacobian>
static void sweepOne(mb2_solver_function* f, const FunctionTables& T, int b, const float* theta, double* errOut, float* stateOut) {
  std::vector<float> jp(size_t(T.numJoints) * 7), js(size_t(T.numJoints) * kJointStateStride), rec(T.recStride + 4);
  for (int row = 0; row < T.numJoints * 7; ++row) jp[row] = jointParameterRow(T, row, theta);
  // the kernels' three passes (fkJoint, the statement-by-statement form, is checked against them below)
  for (int j = 0; j < T.numJoints; ++j) fkLocalFromTheta<kJacobian>(T, j, theta, js.data()); // as sweepKernel: joint parameters straight from theta
  for (int lvl = 1; lvl < T.numLevels; ++lvl)
    for (int k = T.levelStart[lvl]; k < T.levelStart[lvl + 1]; ++k) fkCompose(T, T.levelJoints[k], js.data());
  if (kJacobian)
    for (int i = 0; i < 3 * T.numJoints; ++i) fkAxis(T, i / 3, i % 3, js.data());
  if (kJacobian && getenv("MB2_EMU_CHECK_FK") != nullptr) { // both forms of JointStateT::set agree to rounding
    std::vector<float> ref(js.size());
    for (int lvl = 0; lvl < T.numLevels; ++lvl)
      for (int k = T.levelStart[lvl]; k < T.levelStart[lvl + 1]; ++k) fkJoint<true>(T, T.levelJoints[k], jp.data(), ref.data());
    for (size_t i = 0; i < js.size(); ++i) {
      const float scale = std::max(1.f, std::fabs(ref[i]));
      if (std::fabs(js[i] - ref[i]) > 2e-5f * scale) { std::fprintf(stderr, "emu: three-pass FK differs from fkJoint at %zu: %g vs %g\n", i, js[i], ref[i]); std::abort(); }
    }
  }
  if (stateOut)
    for (int i = 0; i < T.numJoints * 8; ++i) stateOut[i] = js[(i >> 3) * kJointStateStride + (i & 7)];
  const float* tg = f->targets.data() + size_t(b) * T.targetStride;
  const float* cw = f->weights.data() + (T.weightsPerInstance ? size_t(b) * T.numWeights : 0);
  float* Jb = f->J.data() + size_t(b) * T.jacobianStride;
  float* res = kJacobian ? Jb + (T.stripMode ? size_t(T.residOff) : size_t(T.numCols) * T.ldJ) : nullptr;
  // lanes accumulate in double, then a butterfly reduction: emulate the same association
  double lane[32];
  for (int l = 0; l < 32; ++l) lane[l] = 0.0;
  for (int u = 0; u < T.numUnits; ++u) lane[u & 31] += (double)evalUnit<kJacobian>(T, u, theta, (u & 1) ? jp.data() : nullptr, js.data(), tg, cw, rec.data(), res); // both forms of the joint-parameter access
  for (int o = 16; o > 0; o >>= 1) {
    double t[32];
    for (int l = 0; l < 32; ++l) t[l] = lane[l] + lane[l ^ o];
    for (int l = 0; l < 32; ++l) lane[l] = t[l];
  }
  if (kJacobian)
    for (int c = 0; c < T.numCells; ++c) jacobianCell(T, c, js.data(), rec.data(), tg, Jb);
  *errOut = kJacobian ? lane[0] : (double)(float)lane[0];
}

static void jtjOne(const mb2_solver_function* f, int b, int ns, float* H, int ldH) {
  const int nc = f->plan.numCols;
  const float* J = f->J.data() + size_t(b) * (nc + 1) * f->ldJ;
  const float* r = J + size_t(nc) * f->ldJ;
  const int K = roundUp(std::max(f->plan.numRows, 1), 4);
  for (int i = 0; i < ns; ++i) {
    for (int j = 0; j <= i; ++j) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s = fmaf(J[size_t(i) * f->ldJ + k], J[size_t(j) * f->ldJ + k], s);
      H[size_t(j) * ldH + i] = s; // upper triangle only (row j, column i >= j), like the tensor-core epilogue
    }
    float g = 0.f;
    for (int k = 0; k < K; ++k) g = fmaf(J[size_t(i) * f->ldJ + k], r[k], g);
    H[size_t(i) * ldH + ns] = g;
  }
}

// emulation of choleskyKernel<NB> for one instance; returns fail flag, writes delta and g.delta
template <int NB>
static int cholOne(float* Hg, int n, int ldH, float reg, float* delta, float* gdd) {
  const int lda = n | 1;
  std::vector<float> A(size_t(n + 1) * lda, 0.f);
  const int ldp = ((n + 1 + 3) & ~3) + 4;
  std::vector<float> P(size_t(NB) * ldp, 0.f), gsave(n);
  for (int i = 0; i <= n; ++i)
    for (int j = 0; j < n; ++j)
      if (j <= i) A[size_t(i) * lda + j] = Hg[size_t(j) * ldH + i] + ((i == j) ? reg : 0.f);
  for (int i = 0; i < n; ++i) gsave[i] = Hg[size_t(i) * ldH + n];
  int flag = 0;
  CholCtx ctx{A.data(), lda, n, P.data(), ldp, &flag};
  const int blockSize = cholBlockSize(n, NB);
  for (int k = 0; k < n && flag == 0; k += blockSize) {
    const int bs = std::min(blockSize, n - k);
    for (int jj = 0; jj < bs; ++jj) {
      const float x = cholDiagPivot(ctx, k, jj);
      if (!(x > 0.f)) { flag = k + jj + 1; break; }
      for (int lane = 0; lane < 32; ++lane) cholDiagColumn(ctx, k, bs, jj, x, lane);
    }
    if (flag) break;
    for (int t = 0; t < kCholThreads; ++t) cholPanelSolve<NB>(ctx, k, bs, t, kCholThreads);
    for (int t = 0; t < kCholThreads; ++t) cholTrailingUpdate<NB>(ctx, k, bs, t, kCholThreads);
  }
  float* y = A.data() + size_t(n) * lda;
  if (flag) cholForwardFrom(ctx, cholCompletedColumns(flag - 1, blockSize), y);
  for (int i = n - 1; i >= 0; --i) {
    float s = y[i];
    for (int k = i + 1; k < n; ++k) s -= A[size_t(k) * lda + i] * y[k];
    y[i] = s / A[size_t(i) * lda + i];
  }
  float gd = 0.f;
  for (int i = 0; i < n; ++i) { delta[i] = y[i]; gd += gsave[i] * y[i]; }
  *gdd = gd;
  return flag;
}

static int cholDispatch(float* Hg, int n, int ldH, float reg, float* delta, float* gdd) {
  const int eig = cholBlockSize(n, 32);
  if (eig <= 8) return cholOne<8>(Hg, n, ldH, reg, delta, gdd);
  if (eig <= 16) return cholOne<16>(Hg, n, ldH, reg, delta, gdd);
  return cholOne<32>(Hg, n, ldH, reg, delta, gdd);
}

// emulation of gramTilesKernel for one instance: strips = TMA boxes of the K-major Jacobian, warps / half-warps in sequence
static void gramOne(const mb2_solver_function* f, int b, const GramPlan& G, const CholSchedDev& S, float reg, float* out) {
  std::vector<float> store(size_t(G.stride) + 64 + 8, 0.f); // + the all-zero strip
  float* strips = store.data();
  while ((reinterpret_cast<uintptr_t>(strips) & 15) != 0) ++strips;
  std::copy(f->J.data() + size_t(b) * G.stride, f->J.data() + size_t(b + 1) * G.stride, strips); // the bulk copy
  float* resid = strips + G.residOff;
  std::vector<float> tileBuf(256 + 8);
  for (size_t ti = 0; ti < G.tileOrder.size(); ++ti) {
    const int t = G.tileOrder[ti];
    if (t < 0) continue;
    float* tile = out + size_t(t) * 256;
    for (int lane = 0; lane < 32; ++lane) {
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      gramTileAccumulate(strips, G.quad.data(), G.tileQuadStart[t], G.tileQuadStart[t + 1], lane, acc);
      gramTileStore(tile, acc, S.tileInfo[3 * t + 2], reg, lane);
    }
  }
  float* y = out + size_t(G.numTiles) * 256;
  std::vector<int32_t> stripRow(G.numStrips);
  for (int i = 0; i < G.numStrips; ++i) stripRow[i] = G.stripCoord[2 * i];
  for (int K = 0; K < G.numTileCols; ++K)
    for (int hl = 0; hl < 16; ++hl) y[16 * K + hl] = gramVectorEntry(strips, resid, G.colStrip.data(), stripRow.data(), G.colStripStart[K], G.colStripStart[K + 1], hl);
}

// emulation of choleskyScheduledKernel for one instance: phases in the same order, half-warps/warps in sequence
static int cholScheduledOne(const CholSchedDev& S, const float* Hs, int ldH, int n, float reg, float* delta, float* gdd, const float* gramOut = nullptr) {
  std::vector<float> store(size_t(S.numTiles) * 256 + S.nPad + 16, 0.f);
  float* tl = store.data();
  while ((reinterpret_cast<uintptr_t>(tl) & 15) != 0) ++tl; // float4 alignment
  float* y = tl + size_t(S.numTiles) * 256;
  std::vector<float> gsub(n, 0.f);
  std::fill(delta, delta + n, 0.f); // alignment columns keep a zero step
  if (gramOut != nullptr) { // tiles (+ lambda, identity extension) and the slot-ordered J^T r straight from the Gram kernel
    std::copy(gramOut, gramOut + size_t(S.numTiles) * 256 + S.nPad, tl);
    for (int s2 = 0; s2 < S.nPad; ++s2) { const int p = S.perm[s2]; if (p >= 0) gsub[p] = y[s2]; }
  } else {
  for (int t = 0; t < S.numTiles; ++t) { // the TMA box: 16 rows x 16 columns of H starting at (gj0, gi0), zero outside [ns+1] x [ldH]
    const int gi0 = S.tileInfo[3 * t], gj0 = S.tileInfo[3 * t + 1];
    auto boxAt = [&](int row, int col) { return (gj0 + row <= n && gi0 + col < ldH) ? Hs[size_t(gj0 + row) * ldH + gi0 + col] : 0.f; };
    for (int c = 0; c < 16; ++c)
      for (int r = 0; r < 16; ++r)
        tl[size_t(t) * 256 + tileIdx(r, c)] = cholPadElement(boxAt(((S.tileInfo[3 * t + 2] >> 16) & 1) && c > r ? r : c, ((S.tileInfo[3 * t + 2] >> 16) & 1) && c > r ? c : r), S.tileInfo[3 * t + 2], r, c); // cholConvertBox (diagonal boxes mirrored from the upper triangle)
  }
  for (int s2 = 0; s2 < S.nPad; ++s2) {
    const int p = S.perm[s2];
    const float g = p >= 0 ? Hs[size_t(p) * ldH + n] : 0.f;
    y[s2] = g;
    if (p >= 0) { gsub[p] = g; tl[size_t(S.diagTile[s2 >> 4]) * 256 + tileIdx(s2 & 15, s2 & 15)] += reg; }
  }
  }
  int flag = 0;
  for (int L = 0; L < S.numLevels; ++L) {
    for (int ci = S.levelColStart[L]; ci < S.levelColStart[L + 1]; ++ci) {
      const int K = S.levelCols[ci];
      for (int lane = 0; lane < 32; ++lane) cholDiagTile(tl + size_t(S.diagTile[K]) * 256, y + 16 * K, lane, reg, &flag);
    }
    for (int pi = S.levelPanelStart[L]; pi < S.levelPanelStart[L + 1]; ++pi)
    {
      float x[32][2][4];
      for (int lane = 0; lane < 32; ++lane) cholPanelProduct(tl + size_t(S.panelTile[pi]) * 256, tl + size_t(S.panelDiag[pi]) * 256, lane, x[lane]);
      for (int lane = 0; lane < 32; ++lane) cholPanelStore(tl + size_t(S.panelTile[pi]) * 256, lane, x[lane]);
    }
    for (int ti = S.levelTaskStart[L]; ti < S.levelTaskStart[L + 1]; ++ti)
      for (int lane = 0; lane < 32; ++lane) cholUpdateTask(tl, S, ti, lane);
    for (int vi = S.levelVTaskStart[L]; vi < S.levelVTaskStart[L + 1]; ++vi)
      for (int hl = 0; hl < 16; ++hl) cholVectorTask(tl, y, S, vi, hl);
  }
  for (int L = S.numLevels - 1; L >= 0; --L)
    for (int ci = S.levelColStart[L]; ci < S.levelColStart[L + 1]; ++ci)
      for (int lane = 0; lane < 32; ++lane) cholBackwardColumn(tl, y, S, S.levelCols[ci], lane);
  float gd = 0.f;
  for (int i = 0; i < S.nPad; ++i) { const int p = S.perm[i]; if (p >= 0) { delta[p] = y[i]; gd += gsub[p] * y[i]; } }
  *gdd = gd;
  return flag;
}

extern "C" {

const char* mb2_last_error(void) { return g_err.c_str(); }
int mb2_device_count(void) { return 0; }
void mb2_default_gauss_newton_options(mb2_gauss_newton_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->min_iterations = 1; o->max_iterations = 2; o->threshold = 1.f; o->regularization = 0.05f; o->target_rows_per_chunk = ~0ull;
}

int mb2_character_create(int, int32_t J, const int32_t* parents, const float* offsets, const float* prerot, int32_t n, const int32_t* outer,
                         const int32_t* inner, const float* vals, const float* ptoffsets, mb2_character** out) {
  auto c = std::make_unique<mb2_character>();
  HostCharacter& h = c->host;
  h.numJoints = J; h.numParams = n;
  h.parent.assign(parents, parents + J);
  h.offset.assign(offsets, offsets + 3 * J);
  h.prerot.assign(prerot, prerot + 4 * J);
  h.ptOuter.assign(outer, outer + 7 * J + 1);
  const int nnz = outer[7 * J];
  h.ptInner.assign(inner, inner + nnz);
  h.ptVals.assign(vals, vals + nnz);
  h.ptOffsets.assign(ptoffsets, ptoffsets + 7 * J);
  const std::string e = h.validate();
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  h.buildLevels();
  *out = c.release();
  return MB2_OK;
}
int mb2_character_set_parameter_limits(mb2_character* c, int32_t count, const mb2_parameter_limit* limits) {
  c->host.limits.clear();
  for (int i = 0; i < count; ++i) {
    HostLimit l;
    l.type = limits[i].type; l.weight = limits[i].weight;
    std::memcpy(l.i, limits[i].i, sizeof(l.i)); std::memcpy(l.f, limits[i].f, sizeof(l.f));
    c->host.limits.push_back(l);
  }
  return MB2_OK;
}
void mb2_character_destroy(mb2_character* c) { delete c; }

int mb2_solver_function_create(const mb2_character* c, int32_t batch, mb2_solver_function** out) {
  auto f = std::make_unique<mb2_solver_function>();
  f->ch = c; f->B = batch;
  f->enabled.assign(c->host.numParams, 1);
  *out = f.release();
  return MB2_OK;
}
void mb2_solver_function_destroy(mb2_solver_function* f) { delete f; }
int32_t mb2_solver_function_num_parameters(const mb2_solver_function* f) { return f->ch->host.numParams; }
int32_t mb2_solver_function_batch(const mb2_solver_function* f) { return f->B; }
int32_t mb2_solver_function_actual_parameters(const mb2_solver_function* f) {
  int ap = 0;
  for (int i = 0; i < f->ch->host.numParams; ++i) if (f->enabled[i]) ap = i + 1;
  return ap;
}
int32_t mb2_solver_function_jacobian_rows(const mb2_solver_function* f) {
  int total = 0;
  for (const auto& ef : f->efs) if (ef.weight > 0.f) total += jacobianBlockSize(f->ch->host, ef);
  return roundUp(total, 8);
}
int32_t mb2_solver_function_jacobian_stride(const mb2_solver_function* f) {
  int total = 0;
  for (const auto& ef : f->efs) if (ef.weight > 0.f) total += jacobianBlockSize(f->ch->host, ef);
  return std::max(32, roundUp(total, 32));
}

static int addBlock(mb2_solver_function* f, HostErrorFunction& ef, int32_t* outIndex) {
  ef.targetOff = f->targetStride; ef.weightOff = f->numWeights;
  f->targetStride += ef.targetSize;
  if (ef.kind <= 2 || ef.kind == 5) { f->numWeights += ef.numConstraints(); f->weights.insert(f->weights.end(), ef.weights.begin(), ef.weights.end()); }
  f->efs.push_back(ef);
  // re-layout targets (tests add all blocks before setting targets)
  f->targets.assign(size_t(f->B) * std::max(f->targetStride, 1), 0.f);
  if (outIndex) *outIndex = int32_t(f->efs.size()) - 1;
  return MB2_OK;
}
int mb2_add_position_error_function(mb2_solver_function* f, float weight, float alpha, float c, int32_t nc, const int32_t* parents, const float* offsets,
                                    const float* weights, int32_t* outIndex) {
  HostErrorFunction ef;
  ef.kind = 0; ef.weight = weight; ef.lossAlpha = alpha; ef.lossC = c;
  ef.parents.assign(parents, parents + nc); ef.offsets.assign(offsets, offsets + 3 * size_t(nc)); ef.weights.assign(weights, weights + nc);
  ef.targetSize = 3 * nc;
  return addBlock(f, ef, outIndex);
}
int mb2_add_position_error_function_instanced(mb2_solver_function* f, float weight, float alpha, float c, int32_t nc, const int32_t* parents, const float* weights,
                                              int32_t* outIndex) {
  HostErrorFunction ef;
  ef.kind = 0; ef.weight = weight; ef.lossAlpha = alpha; ef.lossC = c; ef.instanceOffsets = true;
  ef.parents.assign(parents, parents + nc);
  ef.offsets.assign(3 * size_t(nc), 0.f);
  ef.weights.assign(weights, weights + nc);
  ef.targetSize = 6 * nc;
  return addBlock(f, ef, outIndex);
}
int mb2_add_plane_error_function(mb2_solver_function* f, float weight, float alpha, float c, int32_t above, int32_t nc, const int32_t* parents,
                                 const float* offsets, const float* weights, int32_t* outIndex) {
  HostErrorFunction ef;
  ef.kind = 5; ef.weight = weight; ef.lossAlpha = alpha; ef.lossC = c; ef.halfPlane = above != 0;
  ef.parents.assign(parents, parents + nc); ef.offsets.assign(offsets, offsets + 3 * size_t(nc)); ef.weights.assign(weights, weights + nc);
  ef.targetSize = 4 * nc;
  return addBlock(f, ef, outIndex);
}
int mb2_add_model_parameters_error_function(mb2_solver_function* f, float weight, const float* targetWeights, int32_t* outIndex) {
  HostErrorFunction ef;
  ef.kind = 6; ef.weight = weight;
  ef.paramWeights.assign(targetWeights, targetWeights + f->ch->host.numParams);
  ef.targetSize = f->ch->host.numParams;
  return addBlock(f, ef, outIndex);
}
int mb2_add_orientation_error_function(mb2_solver_function* f, float weight, float alpha, float c, int32_t rotDiff, int32_t nc, const int32_t* parents,
                                       const float* offsets, const float* weights, int32_t* outIndex) {
  HostErrorFunction ef;
  ef.kind = rotDiff ? 2 : 1; ef.weight = weight; ef.lossAlpha = alpha; ef.lossC = c;
  ef.parents.assign(parents, parents + nc); ef.offsets.assign(offsets, offsets + 4 * size_t(nc));
  for (int i = 0; i < nc; ++i) {
    float* q = &ef.offsets[4 * size_t(i)];
    const float nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) q[k] /= nrm;
  }
  ef.weights.assign(weights, weights + nc);
  ef.targetSize = 4 * nc;
  return addBlock(f, ef, outIndex);
}
int mb2_add_state_error_function(mb2_solver_function* f, float weight, int32_t rotationErrorType, float posWgt, float rotWgt, const float* posW,
                                 const float* rotW, int32_t* outIndex) {
  HostErrorFunction ef;
  ef.kind = 3; ef.weight = weight; ef.rotationErrorType = rotationErrorType; ef.posWgt = posWgt; ef.rotWgt = rotWgt;
  const int J = f->ch->host.numJoints;
  ef.posW.assign(posW, posW + J); ef.rotW.assign(rotW, rotW + J);
  ef.targetSize = 8 * J;
  return addBlock(f, ef, outIndex);
}
int mb2_add_limit_error_function(mb2_solver_function* f, float weight, float alpha, float c, int32_t* outIndex) {
  HostErrorFunction ef;
  ef.kind = 4; ef.weight = weight; ef.lossAlpha = alpha; ef.lossC = c; ef.targetSize = 0;
  return addBlock(f, ef, outIndex);
}
int mb2_set_error_function_weight(mb2_solver_function* f, int32_t index, float weight) { f->efs[index].weight = weight; return MB2_OK; }
int mb2_set_targets(mb2_solver_function* f, int32_t index, const float* targets) {
  const HostErrorFunction& ef = f->efs[index];
  for (int b = 0; b < f->B; ++b) {
    float* dst = f->targets.data() + size_t(b) * f->targetStride + ef.targetOff;
    std::memcpy(dst, targets + size_t(b) * ef.targetSize, size_t(ef.targetSize) * sizeof(float));
    if (ef.kind == 1 || ef.kind == 2)
      for (int q = 0; q < ef.numConstraints(); ++q) {
        float* p = dst + 4 * q;
        const float n = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
        p[0] /= n; p[1] /= n; p[2] /= n; p[3] /= n;
      }
  }
  return MB2_OK;
}
int mb2_set_constraint_weights(mb2_solver_function* f, int32_t index, const float* weights, int32_t perInstance) {
  HostErrorFunction& ef = f->efs[index];
  const int nc = ef.numConstraints();
  if (!perInstance) { std::copy(weights, weights + nc, f->weights.begin() + ef.weightOff); return MB2_OK; }
  if (!f->weightsPerInstance) {
    std::vector<float> all(size_t(f->B) * f->numWeights);
    for (int b = 0; b < f->B; ++b) std::copy(f->weights.begin(), f->weights.begin() + f->numWeights, all.begin() + size_t(b) * f->numWeights);
    f->weights = all;
    f->weightsPerInstance = true;
  }
  for (int b = 0; b < f->B; ++b) std::copy(weights + size_t(b) * nc, weights + size_t(b + 1) * nc, f->weights.begin() + size_t(b) * f->numWeights + ef.weightOff);
  return MB2_OK;
}
int mb2_solver_function_set_enabled_parameters(mb2_solver_function* f, const uint64_t* bits) {
  for (int i = 0; i < f->ch->host.numParams; ++i) f->enabled[i] = (bits[i >> 6] >> (i & 63)) & 1ull ? 1 : 0;
  return MB2_OK;
}

int mb2_solver_function_get_error(mb2_solver_function* f, const float* params, double* errors) {
  const std::string e = plan(f, f->planCompact);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  const FunctionTables T = tables(f);
  for (int b = 0; b < f->B; ++b) sweepOne<false>(f, T, b, params + size_t(b) * T.numParams, &errors[b], nullptr);
  return MB2_OK;
}
int mb2_solver_function_get_skeleton_state(mb2_solver_function* f, const float* params, float* state) {
  const std::string e = plan(f, f->planCompact);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  const FunctionTables T = tables(f);
  double err;
  for (int b = 0; b < f->B; ++b) sweepOne<false>(f, T, b, params + size_t(b) * T.numParams, &err, state + size_t(b) * T.numJoints * 8);
  return MB2_OK;
}
int mb2_solver_function_get_jacobian(mb2_solver_function* f, const float* params, float* jac, float* residual, double* errors, int32_t* actualRows) {
  const std::string e = plan(f, false);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  const FunctionTables T = tables(f);
  const int rows = mb2_solver_function_jacobian_rows(f), n = T.numParams;
  for (int b = 0; b < f->B; ++b) {
    double err;
    sweepOne<true>(f, T, b, params + size_t(b) * n, &err, nullptr);
    if (errors) errors[b] = err;
    for (int c = 0; c < n && jac; ++c) std::memcpy(jac + (size_t(b) * n + c) * rows, f->J.data() + (size_t(b) * (n + 1) + c) * f->ldJ, size_t(rows) * sizeof(float));
    if (residual) std::memcpy(residual + size_t(b) * rows, f->J.data() + (size_t(b) * (n + 1) + n) * f->ldJ, size_t(rows) * sizeof(float));
  }
  if (actualRows) *actualRows = rows;
  return MB2_OK;
}
int mb2_solver_function_get_jtjr(mb2_solver_function* f, const float* params, int32_t, float* jtj, float* jtr, double* errors) {
  const std::string e = plan(f, false);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  const FunctionTables T = tables(f);
  const int ap = f->plan.actualParameters, n = T.numParams, ldH = roundUp(ap + 1, 16);
  std::vector<float> H(size_t(ap + 1) * ldH);
  for (int b = 0; b < f->B; ++b) {
    double err;
    sweepOne<true>(f, T, b, params + size_t(b) * n, &err, nullptr);
    if (errors) errors[b] = err;
    std::fill(H.begin(), H.end(), 0.f);
    jtjOne(f, b, ap, H.data(), ldH);
    for (int j = 0; j < ap; ++j) {
      if (jtj) for (int i = j; i < ap; ++i) jtj[(size_t(b) * ap + i) * ap + j] = H[size_t(j) * ldH + i];
      if (jtr) jtr[size_t(b) * ap + j] = H[size_t(j) * ldH + ap];
    }
  }
  return MB2_OK;
}

int mb2_solver_create(mb2_solver_function* f, const mb2_gauss_newton_options* opt, mb2_solver** out) {
  auto s = std::make_unique<mb2_solver>();
  s->fn = f;
  if (opt) s->opt = *opt; else mb2_default_gauss_newton_options(&s->opt);
  *out = s.release();
  return MB2_OK;
}
void mb2_solver_destroy(mb2_solver* s) { delete s; }
int mb2_solver_set_options(mb2_solver* s, const mb2_gauss_newton_options* opt) { s->opt = *opt; return MB2_OK; }
int mb2_solver_set_enabled_parameters(mb2_solver* s, const uint64_t* bits) { return mb2_solver_function_set_enabled_parameters(s->fn, bits); }

// emulation of mb2_solver_solve_device's launch sequence, instance by instance
int mb2_solver_solve(mb2_solver* s, float* params, double* errors, int32_t* iterations, int32_t* status) {
  mb2_solver_function* f = s->fn;
  const auto& o = s->opt;
  int numEnabled = 0;
  for (uint8_t e : f->enabled) numEnabled += e ? 1 : 0;
  int cholMode = o.cholesky_mode;
  if (cholMode == 0) cholMode = numEnabled >= 48 ? 3 : 1;
  std::string e = plan(f, true);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  CholSchedule sched;
  GramPlan gram;
  const bool useGram = cholMode >= 2 && (o.jtj_mode == MB2_JTJ_AUTO || o.jtj_mode == MB2_JTJ_SPARSE_TILES);
  if (cholMode >= 2) { // same two-pass planning as ensurePlan(mode 2) in ik_capi.cu
    const int ns0 = f->plan.numCols;
    std::vector<std::vector<int>> cliques(f->plan.units.size());
    for (const CellDesc& c : f->plan.cells) cliques[c.unit].push_back(int(c.col));
    const std::vector<int> prio = columnDepthPriority(f->ch->host, f->plan.enabledList);
    e = buildCholSchedule(ns0, cliques, cholMode == 2, sched, &prio);
    if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
    std::vector<int32_t> colOrder;
    layoutDeviceColumns(sched, colOrder);
    for (int32_t& c : colOrder) if (c >= 0) c = f->plan.enabledList[c];
    e = buildPlan(f->ch->host, f->efs, f->enabled, true, f->plan, &colOrder, useGram);
    if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
    if (useGram) {
      std::vector<int32_t> cr0, crn, cc;
      for (const CellDesc& c : f->plan.cells) { cr0.push_back(f->plan.units[c.unit].row0); crn.push_back(f->plan.units[c.unit].numRows); cc.push_back(int32_t(c.col)); }
      e = buildGramPlan(sched, cr0, crn, cc, f->plan.numRows, gram);
      if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
      for (size_t i = 0; i < f->plan.cells.size(); ++i) { f->plan.cells[i].stripOff = gram.cellStripOff[i]; f->plan.cells[i].quadStride = gram.cellQuadStride[i]; }
    }
    f->stripMode = useGram; f->residOff = gram.residOff; f->stripStride = gram.stride;
    f->ldJ = std::max(32, roundUp(f->plan.numRows, 32));
    f->J.assign(useGram ? size_t(f->B) * gram.stride : size_t(f->B) * (f->plan.numCols + 1) * f->ldJ, 0.f);
  }
  std::vector<int32_t> blob;
  CholSchedDev S{};
  if (cholMode >= 2) makeScheduleBlob(sched, blob, S);
  const FunctionTables T = tables(f);
  const int n = T.numParams, ns = f->plan.numCols, ldH = roundUp(ns + 1, 16);
  const int maxIt = int(o.max_iterations), minIt = int(o.min_iterations);
  s->errors.assign(f->B, DBL_MAX); s->iterations.assign(f->B, 0); s->status.assign(f->B, 0);
  s->history.assign(size_t(f->B) * std::max(maxIt, 1), 0.0);
  std::vector<float> H(size_t(ns + 1) * ldH, 0.f), delta(ns), orig(n), gramOut(useGram ? size_t(sched.numTiles) * 256 + sched.nPad : 1, 0.f);
  s->totalIterations = 0;
  for (int b = 0; b < f->B; ++b) {
    float* theta = params + size_t(b) * n;
    std::vector<float> theta0(theta, theta + n);
    double last = DBL_MAX, error = DBL_MAX;
    for (int it = 0; it < maxIt; ++it) {
      sweepOne<true>(f, T, b, theta, &error, nullptr);
      float gdd = 0.f;
      int failed;
      std::fill(H.begin(), H.end(), std::nanf("")); // entries the device never writes (lower triangle) are garbage there: poison them here
      jtjOne(f, b, ns, H.data(), ldH);
      if (cholMode >= 2 && useGram) {
        gramOne(f, b, gram, S, o.regularization, gramOut.data());
        failed = cholScheduledOne(S, nullptr, ldH, ns, o.regularization, delta.data(), &gdd, gramOut.data());
      } else if (cholMode >= 2) {
        failed = cholScheduledOne(S, H.data(), ldH, ns, o.regularization, delta.data(), &gdd);
      } else {
        failed = cholDispatch(H.data(), ns, ldH, o.regularization, delta.data(), &gdd);
      }
      if (failed && s->status[b] == 0) s->status[b] = MB2_INSTANCE_CHOLESKY_BREAKDOWN;
      if (!o.do_line_search) {
        for (int a = 0; a < ns; ++a) if (f->plan.deviceCols[a] >= 0) theta[f->plan.deviceCols[a]] -= delta[a];
      } else {
        std::copy(theta, theta + n, orig.begin());
        float scale = 1.f;
        for (int step = 0; step < 10; ++step) {
          for (int a = 0; a < ns; ++a) { const int c = f->plan.deviceCols[a]; if (c >= 0) theta[c] = orig[c] - scale * delta[a]; }
          double en;
          sweepOne<false>(f, T, b, theta, &en, nullptr);
          bool accept;
          if (!o.subset_line_search) accept = (error - en) >= (double)(scale * (1e-3f * (float)error));
          else accept = (error - en) >= (double)(1e-4f * scale) * (double)gdd;
          if (accept || step >= 9) break;
          scale *= 0.5f;
        }
      }
      s->history[size_t(b) * std::max(maxIt, 1) + it] = error;
      s->iterations[b] = it + 1;
      const bool converged = std::fabs(last - error) / (std::fabs(error) + (double)FLT_MIN) <= (double)(o.threshold * FLT_EPSILON);
      if (it >= minIt && converged) break;
      last = error;
    }
    bool bad = false;
    for (int i = 0; i < n; ++i) if (!std::isfinite(theta[i])) bad = true;
    if (bad) { std::copy(theta0.begin(), theta0.end(), theta); s->status[b] = MB2_INSTANCE_NON_FINITE; }
    s->errors[b] = error;
    s->totalIterations += s->iterations[b];
    if (errors) errors[b] = error;
    if (iterations) iterations[b] = s->iterations[b];
    if (status) status[b] = s->status[b];
  }
  return MB2_OK;
}
int mb2_solver_get_error_history(mb2_solver* s, double* history) {
  std::memcpy(history, s->history.data(), s->history.size() * sizeof(double));
  return MB2_OK;
}
int mb2_solver_get_counters(mb2_solver* s, uint64_t* totalIterations, uint64_t* kernelLaunches) {
  if (totalIterations) *totalIterations = s->totalIterations;
  if (kernelLaunches) *kernelLaunches = 0;
  return MB2_OK;
}

} // extern "C"

// ---- scheduler statistics (test/debug helper) ----
#include "../../momentum_b200/csrc/ik_chol_sched.h"
extern "C" int emu_chol_schedule_stats(int n, int numCliques, const int* cliqueStart, const int* cliqueCols, int forceDense, long long* stats) {
  std::vector<std::vector<int>> cl(numCliques);
  for (int c = 0; c < numCliques; ++c) cl[c].assign(cliqueCols + cliqueStart[c], cliqueCols + cliqueStart[c + 1]);
  CholSchedule s;
  const std::string e = buildCholSchedule(n, cl, forceDense != 0, s);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  stats[0] = s.numTileCols; stats[1] = s.numTiles; stats[2] = s.numLevels; stats[3] = s.tileOps; stats[4] = s.denseTileOps;
  stats[5] = (long long)s.taskDst.size(); stats[6] = (long long)s.panelTile.size();
  return MB2_OK;
}

extern "C" int emu_chol_schedule_dump(int n, int numCliques, const int* cliqueStart, const int* cliqueCols, int forceDense) {
  std::vector<std::vector<int>> cl(numCliques);
  for (int c = 0; c < numCliques; ++c) cl[c].assign(cliqueCols + cliqueStart[c], cliqueCols + cliqueStart[c + 1]);
  CholSchedule s;
  const std::string e = buildCholSchedule(n, cl, forceDense != 0, s);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  std::printf("n %d nPad %d tilecols %d tiles %d levels %d\n", s.n, s.nPad, s.numTileCols, s.numTiles, s.numLevels);
  for (int L = 0; L < s.numLevels; ++L) {
    int maxPairs = 0, totPairs = 0;
    for (int t = s.levelTaskStart[L]; t < s.levelTaskStart[L + 1]; ++t) { const int p = s.taskPairStart[t + 1] - s.taskPairStart[t]; maxPairs = std::max(maxPairs, p); totPairs += p; }
    std::printf(" level %d: cols %d panels %d tasks %d pairs %d maxPairsPerTask %d vtasks %d | cols:", L, s.levelColStart[L + 1] - s.levelColStart[L],
                s.levelPanelStart[L + 1] - s.levelPanelStart[L], s.levelTaskStart[L + 1] - s.levelTaskStart[L], totPairs, maxPairs,
                s.levelVTaskStart[L + 1] - s.levelVTaskStart[L]);
    for (int c = s.levelColStart[L]; c < s.levelColStart[L + 1]; ++c) {
      int valid = 0;
      for (int r = 0; r < 16; ++r) valid += s.perm[16 * s.levelCols[c] + r] >= 0;
      std::printf(" %d(%d)", s.levelCols[c], valid);
    }
    std::printf("\n");
  }
  return MB2_OK;
}

extern "C" int emu_chol_chunk_stats(int n, int numCliques, const int* cliqueStart, const int* cliqueCols) {
  std::vector<std::vector<int>> cl(numCliques);
  for (int c = 0; c < numCliques; ++c) cl[c].assign(cliqueCols + cliqueStart[c], cliqueCols + cliqueStart[c + 1]);
  CholSchedule s;
  const std::string e = buildCholSchedule(n, cl, false, s);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  std::vector<int32_t> order;
  layoutDeviceColumns(s, order);
  const int nd = s.n + 1; // + residual column
  const int nb = (nd + 31) / 32;
  std::vector<int> need(nb * nb, 0);
  auto validOf = [&](int K) { int v = 0; while (v < 16 && s.perm[16 * K + v] >= 0) ++v; return v; };
  for (int t = 0; t < s.numTiles; ++t) {
    const int I = s.tileRow[t], J = s.tileCol[t];
    const int vI = validOf(I), vJ = validOf(J), gi0 = s.perm[16 * I], gj0 = s.perm[16 * J];
    for (int r = gj0; r < gj0 + vJ; ++r) for (int c = gi0; c < gi0 + vI; ++c) if (c >= r) need[(r / 32) * nb + c / 32] = 1;
  }
  int total = 0, needed = 0;
  for (int a = 0; a < nb; ++a) for (int b2 = a; b2 < nb; ++b2) { ++total; needed += need[a * nb + b2] || b2 == (nd - 1) / 32; }
  std::printf("device columns %d (+1), 32x32 chunks in the upper triangle %d, needed %d\n", s.n, total, needed);
  for (int a = 0; a < nb; ++a) { for (int b2 = 0; b2 < nb; ++b2) std::printf("%c", b2 < a ? ' ' : (need[a * nb + b2] ? '#' : (b2 == (nd - 1) / 32 ? 'g' : '.'))); std::printf("\n"); }
  return MB2_OK;
}

extern "C" int emu_gram_stats(mb2_solver_function* f) {
  std::string e = plan(f, true);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  CholSchedule sched;
  std::vector<std::vector<int>> cliques(f->plan.units.size());
  for (const CellDesc& c : f->plan.cells) cliques[c.unit].push_back(int(c.col));
  const std::vector<int> prio = columnDepthPriority(f->ch->host, f->plan.enabledList);
  e = buildCholSchedule(f->plan.numCols, cliques, false, sched, &prio);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  std::vector<int32_t> colOrder;
  layoutDeviceColumns(sched, colOrder);
  for (int32_t& c : colOrder) if (c >= 0) c = f->plan.enabledList[c];
  e = buildPlan(f->ch->host, f->efs, f->enabled, true, f->plan, &colOrder, true);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  std::vector<int32_t> cr0, crn, cc;
  for (const CellDesc& c : f->plan.cells) { cr0.push_back(f->plan.units[c.unit].row0); crn.push_back(f->plan.units[c.unit].numRows); cc.push_back(int32_t(c.col)); }
  GramPlan g;
  e = buildGramPlan(sched, cr0, crn, cc, f->plan.numRows, g);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  int maxPairs = 0;
  for (int t = 0; t < g.numTiles; ++t) maxPairs = std::max(maxPairs, g.tilePairStart[t + 1] - g.tilePairStart[t]);
  std::printf("rows %d (aligned), device columns %d, strips %d (%d KB), tiles %d, pairs %zu (max per tile %d), MACs %lld\n", f->plan.numRows, f->plan.numCols, g.numStrips,
              g.numStrips / 4, g.numTiles, g.pairA.size(), maxPairs, (long long)g.macs);
  std::printf("pairs per tile in order:");
  for (size_t ti = 0; ti < g.tileOrder.size(); ++ti) { const int t = g.tileOrder[ti]; if (t >= 0) std::printf(" %d", g.tilePairStart[t + 1] - g.tilePairStart[t]); else std::printf(" -"); }
  std::printf("\n");
  return MB2_OK;
}

extern "C" int emu_sched_structure(mb2_solver_function* f) {
  std::string e = plan(f, true);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  CholSchedule s;
  std::vector<std::vector<int>> cliques(f->plan.units.size());
  for (const CellDesc& c : f->plan.cells) cliques[c.unit].push_back(int(c.col));
  const std::vector<int> prio = columnDepthPriority(f->ch->host, f->plan.enabledList);
  e = buildCholSchedule(f->plan.numCols, cliques, false, s, &prio);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  std::printf("levels %d tiles %d\n", s.numLevels, s.numTiles);
  for (int L = 0; L < s.numLevels; ++L) { std::printf(" level %d:", L); for (int c = s.levelColStart[L]; c < s.levelColStart[L + 1]; ++c) std::printf(" %d", s.levelCols[c]); std::printf("\n"); }
  const int T = s.numTileCols;
  const HostCharacter& h = f->ch->host;
  // parameter -> a joint it drives (first one)
  std::vector<int> jointOf(h.numParams, -1);
  for (int r = 0; r < 7 * h.numJoints; ++r)
    for (int k = h.ptOuter[r]; k < h.ptOuter[r + 1]; ++k) if (jointOf[h.ptInner[k]] < 0) jointOf[h.ptInner[k]] = r / 7;
  for (int K = 0; K < T; ++K) {
    std::printf("tile col %2d: joints {", K);
    int last = -2;
    for (int j = 0; j < 16; ++j) { const int p = s.perm[16 * K + j]; if (p < 0) continue; const int jt = jointOf[f->plan.enabledList[p]]; if (jt != last) std::printf(" %d", jt); last = jt; }
    std::printf(" }  rows below:");
    for (int I = K + 1; I < T; ++I) if (s.tileIdTable[size_t(I) * T + K] >= 0) std::printf(" %d", I);
    std::printf("\n");
  }
  std::printf("parents:");
  for (int j = 0; j < h.numJoints; ++j) std::printf(" %d:%d", j, h.parent[j]);
  std::printf("\n");
  return MB2_OK;
}

// schedule / Gram-plan figures of the solver plan (mode 2 planning of ik_capi.cu): out = {levels, tiles, tile columns, nPad, device
// columns, strips, pairs, misaligned tile starts, odd pair lists, cells outside their strip}
extern "C" int emu_plan_figures(mb2_solver_function* f, int64_t out[10]) {
  std::string e = plan(f, true);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  CholSchedule s;
  std::vector<std::vector<int>> cliques(f->plan.units.size());
  for (const CellDesc& c : f->plan.cells) cliques[c.unit].push_back(int(c.col));
  const std::vector<int> prio = columnDepthPriority(f->ch->host, f->plan.enabledList);
  e = buildCholSchedule(f->plan.numCols, cliques, false, s, &prio);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  std::vector<int32_t> colOrder;
  layoutDeviceColumns(s, colOrder);
  for (int32_t& c : colOrder) if (c >= 0) c = f->plan.enabledList[c];
  e = buildPlan(f->ch->host, f->efs, f->enabled, true, f->plan, &colOrder, true);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  std::vector<int32_t> cr0, crn, cc;
  for (const CellDesc& c : f->plan.cells) { cr0.push_back(f->plan.units[c.unit].row0); crn.push_back(f->plan.units[c.unit].numRows); cc.push_back(int32_t(c.col)); }
  GramPlan g;
  e = buildGramPlan(s, cr0, crn, cc, f->plan.numRows, g);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  int64_t misaligned = 0, odd = 0, outside = 0;
  for (int K = 0; K < s.numTileCols; ++K) misaligned += (s.perm[16 * K] < 0 || (s.perm[16 * K] & 3) != 0) ? 1 : 0;
  for (int t = 0; t < s.numTiles; ++t) odd += (g.tilePairStart[t + 1] - g.tilePairStart[t]) & 1;
  for (size_t i = 0; i < cc.size(); ++i) {
    const int strip = int(g.cellStripOff[i] / 64), colInStrip = int(g.cellStripOff[i] % 64) / 4;
    if (strip >= g.numStrips || g.stripCoord[2 * strip] != (cr0[i] & ~3) || g.stripCoord[2 * strip + 1] + colInStrip != cc[i]) ++outside;
  }
  out[0] = s.numLevels; out[1] = s.numTiles; out[2] = s.numTileCols; out[3] = s.nPad; out[4] = s.n; out[5] = g.numStrips; out[6] = int64_t(g.pairA.size());
  out[7] = misaligned; out[8] = odd; out[9] = outside;
  return MB2_OK;
}

// What the sweep kernel's 16-byte strip stores rely on (ik_device.cuh jacobianCell): in the solver (strip) layout every Position /
// Orientation unit starts on a row quad and owns whole quads (the rows up to the next multiple of four belong to no other unit), its
// cells sit at 16-byte aligned strip offsets, and cells are ordered by (kind, unit, device column).
// out: [0] units starting off a quad boundary, [1] units whose padding rows overlap another unit, [2] cells at unaligned strip offsets,
//      [3] cells out of (kind, unit, column) order, [4] multi-row units, [5] cells
extern "C" int emu_store_invariants(mb2_solver_function* f, int64_t out[6]) {
  int64_t fig[10];
  if (emu_plan_figures(f, fig) != MB2_OK) return MB2_ERR_INVALID_ARGUMENT; // leaves f->plan in the solver layout (aligned row groups)
  const Plan& p = f->plan;
  std::vector<int> owner(size_t(p.numRows) + 8, -1);
  int64_t offQuad = 0, overlap = 0, multi = 0;
  for (size_t u = 0; u < p.units.size(); ++u) {
    const UnitDesc& d = p.units[u];
    const bool joint = d.kind == kUnitPosition || d.kind == kUnitOrientation || d.kind == kUnitOrientationRotDiff;
    if (joint) { ++multi; if ((d.row0 & 3) != 0) ++offQuad; }
    const int end = joint ? ((d.row0 + d.numRows + 3) & ~3) : d.row0 + d.numRows; // a joint unit claims its padding rows too
    for (int r = d.row0; r < end; ++r) { if (owner[size_t(r)] >= 0) ++overlap; owner[size_t(r)] = int(u); }
  }
  // strip offsets of the cells: the Gram plan of the same layout (as emu_plan_figures builds it)
  CholSchedule s;
  std::vector<std::vector<int>> cliques(p.units.size());
  for (const CellDesc& c : p.cells) cliques[c.unit].push_back(int(c.col));
  std::vector<int32_t> cr0, crn, cc;
  for (const CellDesc& c : p.cells) { cr0.push_back(p.units[c.unit].row0); crn.push_back(p.units[c.unit].numRows); cc.push_back(int32_t(c.col)); }
  GramPlan g;
  {
    // the schedule must be the one the plan's device columns were laid out for: rebuild it from the natural-order plan
    mb2_solver_function tmp = *f; // (plain copy of the host description; device-free harness)
    std::string e = plan(&tmp, true);
    if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
    std::vector<std::vector<int>> cl(tmp.plan.units.size());
    for (const CellDesc& c : tmp.plan.cells) cl[c.unit].push_back(int(c.col));
    const std::vector<int> prio = columnDepthPriority(f->ch->host, tmp.plan.enabledList);
    e = buildCholSchedule(tmp.plan.numCols, cl, false, s, &prio);
    if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
    std::vector<int32_t> colOrder;
    layoutDeviceColumns(s, colOrder);
    e = buildGramPlan(s, cr0, crn, cc, p.numRows, g);
    if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  }
  int64_t unaligned = 0, disorder = 0;
  for (size_t i = 0; i < p.cells.size(); ++i) {
    const CellDesc& c = p.cells[i];
    if ((g.cellStripOff[i] & 3u) != 0) ++unaligned;
    if (i > 0) {
      const CellDesc& a = p.cells[i - 1];
      const int ka = p.units[a.unit].kind, kb = p.units[c.unit].kind;
      const bool ok = ka < kb || (ka == kb && (a.unit < c.unit || (a.unit == c.unit && a.col < c.col)));
      if (!ok) ++disorder;
    }
  }
  out[0] = offQuad; out[1] = overlap; out[2] = unaligned; out[3] = disorder; out[4] = multi; out[5] = int64_t(p.cells.size());
  return MB2_OK;
}

// table sizes of the solver plan (bytes): what a kernel that stages every table in shared memory has to hold
extern "C" int emu_table_sizes(mb2_solver_function* f, int64_t out[12]) {
  int64_t fig[10];
  if (emu_plan_figures(f, fig) != MB2_OK) return MB2_ERR_INVALID_ARGUMENT; // leaves f->plan in the solver layout
  const HostCharacter& h = f->ch->host;
  out[0] = int64_t(f->plan.units.size());
  out[1] = int64_t(f->plan.cells.size());
  out[2] = int64_t(f->plan.contribs.size());
  out[3] = int64_t(h.ptInner.size());
  out[4] = int64_t(f->plan.limitData.size());
  out[5] = f->plan.recStride;
  out[6] = f->targetStride;
  out[7] = int64_t(sizeof(UnitDesc));
  out[8] = int64_t(sizeof(CellDesc));
  out[9] = int64_t(sizeof(ContribDesc));
  out[10] = int64_t(h.levelStart.size()) - 1;
  int maxContrib = 0;
  for (const CellDesc& c : f->plan.cells) maxContrib = std::max<int>(maxContrib, c.contribCount);
  out[11] = maxContrib;
  return MB2_OK;
}

// per-level update-task pair counts of the solver plan's schedule (balance analysis)
extern "C" int emu_task_loads(mb2_solver_function* f) {
  std::string e = plan(f, true);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  CholSchedule s;
  std::vector<std::vector<int>> cliques(f->plan.units.size());
  for (const CellDesc& c : f->plan.cells) cliques[c.unit].push_back(int(c.col));
  const std::vector<int> prio = columnDepthPriority(f->ch->host, f->plan.enabledList);
  e = buildCholSchedule(f->plan.numCols, cliques, false, s, &prio);
  if (!e.empty()) return fail(MB2_ERR_INVALID_ARGUMENT, e);
  for (int L = 0; L < s.numLevels; ++L) {
    std::printf("level %d: diag %d panels %d tasks:", L, s.levelColStart[L + 1] - s.levelColStart[L], s.levelPanelStart[L + 1] - s.levelPanelStart[L]);
    for (int t = s.levelTaskStart[L]; t < s.levelTaskStart[L + 1]; ++t) std::printf(" %d", s.taskPairStart[t + 1] - s.taskPairStart[t]);
    std::printf(" | vtasks:");
    for (int t = s.levelVTaskStart[L]; t < s.levelVTaskStart[L + 1]; ++t) std::printf(" %d", s.vtaskSrcStart[t + 1] - s.vtaskSrcStart[t]);
    std::printf("\n");
  }
  return MB2_OK;
}

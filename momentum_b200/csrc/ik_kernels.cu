// sm_100a kernels for one Gauss-Newton iteration over a batch of IK instances.
//
//   K1  sweepKernel<true, W>      FK sweep + residual + Jacobian cells (strip layout or K-major matrix)  (skeleton_solver_function.cpp:200-261)
//   K4  sweepKernel<false, W>     FK sweep + error only (line search)                                    (skeleton_solver_function.cpp:64-83)
//   K2s gramTilesKernel           stored tiles of JtJ + lambda I and Jtr from the non-zero strips of J (mma.sync, three-term TF32 split)
//   K2' jtjSimtKernel             dense JtJ and Jtr, fp32 CUDA cores (solver_function.cpp:113-116): validation path / wide systems
//   K3  choleskyScheduledKernel   level-scheduled tile-sparse damped Cholesky + solves + theta -= delta + SolverT bookkeeping
//   K3' choleskyKernel<NB>        dense blocked LLT with Eigen's block structure and early exit (gauss_newton_solver.cpp:248-259, solver.cpp:89-122)
//   small kernels                 line search, bookkeeping, target scatter / quaternion normalisation
//
// The dense tensor-core JtJ (tcgen05, TMEM accumulators, TMA-fed) lives in ik_jtj_tc.cu; PTX wrappers in ik_ptx.cuh.
#include "ik_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <type_traits>

#include "ik_chol.cuh"
#include "ik_chol_sched.h"
#include "ik_jtj_tc.cuh"
#include "ik_ptx.cuh"
#include "ik_device.cuh"

namespace mb2 {

// ------------------------------------------------------------------------------------------------
// K1 / K4: one warp per IK instance. Joint state lives in shared memory (17 floats / joint, odd
// stride => conflict-free when lanes own different joints); the joint tree is swept level by level
// with lanes = joints of one depth level; then lanes = units (constraints), then lanes = Jacobian cells.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warpSum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

size_t sweepSmemPerInstance(const FunctionTables& T, int warpsPerInstance) {
  const size_t nPad = (T.numParams + 3) & ~3;
  // several warps per instance (large rigs): the joint parameters [7 J] are staged too (lanes = rows: the lanes are plentiful there)
  const size_t jp = warpsPerInstance > 1 ? size_t((T.numJoints * kParametersPerJoint + 1) & ~1) : 0;
  return sizeof(float) * (nPad + jp + size_t((T.numJoints * kJointStateStride + 1) & ~1) + size_t((T.recStride + 1) & ~1) + 4 + 2 * size_t(warpsPerInstance));
}

// The read-only tables (character + plan) are walked by dependent loads (cell -> unit -> contributions -> joint);
// from L2 each hop costs several hundred cycles, so a persistent CTA copies them into shared memory once.
MB2_HD size_t tableWords(size_t count, size_t elemBytes) { return (count * elemBytes + 15) / 16 * 4; }
// mode 1: every table; mode 2: everything except the two big ones, cells and contributions (large rigs: those are streamed through
// L1 / L2, one sequential record per lane and iteration, while the tables walked by dependent loads - character, units, error functions -
// still sit in shared memory)
size_t sweepTableBytes(const FunctionTables& T, int mode) {
  if (mode == 0) return 0;
  const size_t J = T.numJoints;
  size_t w = 0;
  w += tableWords(J, 4) + tableWords(3 * J, 4) + tableWords(4 * J, 4);                       // parent, offset, prerot
  w += tableWords(7 * J + 1, 4) + tableWords(T.ptNnz, 4) * 2 + tableWords(7 * J, 4);         // ptOuter, ptInner, ptVals, ptOffsets
  w += tableWords(T.numLevels + 1, 4) + tableWords(J, 4);                                    // levelStart, levelJoints
  w += tableWords(T.numEf, sizeof(EfDesc)) + tableWords(T.numUnits, sizeof(UnitDesc)) + tableWords(T.numLimitData, 4);
  if (mode == 1) w += tableWords(T.numCells, sizeof(CellDesc)) + tableWords(T.numContribs, sizeof(ContribDesc));
  return w * 4;
}

template <class E>
__device__ __forceinline__ void stageTable(const E*& table, size_t count, uint32_t*& cursor) {
  static_assert(sizeof(E) % 4 == 0, "tables are staged word by word");
  const uint32_t* src = reinterpret_cast<const uint32_t*>(table);
  const size_t words = count * (sizeof(E) / 4);
  // eight loads in flight per thread (a word-by-word loop is a chain of global-load latencies: 23 of them for the 40 KB of cfg3)
  const int nt = int(blockDim.x);
  size_t i = threadIdx.x;
  for (; i + 7 * size_t(nt) < words; i += 8 * size_t(nt)) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = src[i + size_t(k) * nt];
#pragma unroll
    for (int k = 0; k < 8; ++k) cursor[i + size_t(k) * nt] = v[k];
  }
  for (; i < words; i += nt) cursor[i] = src[i];
  table = reinterpret_cast<const E*>(cursor);
  cursor += tableWords(count, sizeof(E));
}

// kStage (1 all tables, 2 all but cells / contributions, 0 none): the tables live in shared memory for the whole kernel. A template parameter rather than a run-time branch so that every
// table pointer is PROVABLY a shared-memory address: with `if (a.stageTables)` the pointers could be either and all 273 table loads
// of the kernel were generic LD instructions (address-space check on every hop of the dependent chains cell -> unit -> record ->
// contributions; 43 % of the warp-state samples were long-scoreboard waits on them, profiles/r02k).
constexpr int kSweepMaxWarps = 24; // up to 24 warps per CTA (85 registers per thread)
template <bool kJacobian, int W, int kStage>
__global__ void __launch_bounds__(32 * kSweepMaxWarps) sweepKernel(const SweepArgs a) {
  extern __shared__ __align__(16) float smem[];
  FunctionTables T = a.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a group of W warps works on one instance: gl = lane within the group, gs = lanes of the group
  const int group = warp / W, gl = (warp % W) * 32 + lane;
  constexpr int gs = 32 * W;
  const int groupsPerCta = (blockDim.x >> 5) / W;
  const int nPad = (T.numParams + 3) & ~3;
  // (kJointStateStride is odd: an odd joint count gets one float of padding so that the doubles below stay 8-byte aligned)
  const int jpFloats = W > 1 ? ((T.numJoints * kParametersPerJoint + 1) & ~1) : 0;
  const int perGroup = nPad + jpFloats + ((T.numJoints * kJointStateStride + 1) & ~1) + ((T.recStride + 1) & ~1) + 4 + 2 * W;
  float* th = smem + size_t(group) * perGroup;
  float* jp = W > 1 ? th + nPad : nullptr;
  float* js = th + nPad + jpFloats;
  float* rec = js + ((T.numJoints * kJointStateStride + 1) & ~1);
  double* errSlots = reinterpret_cast<double*>(rec + ((T.recStride + 1) & ~1)); // W partial sums (8-byte aligned: every preceding size is even)
  auto groupSync = [&]() {
    if (W == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(gs) : "memory");
  };
  if constexpr (kStage != 0) {
    uint32_t* cursor = reinterpret_cast<uint32_t*>(smem + ((size_t(groupsPerCta) * perGroup + 3) & ~size_t(3)));
    const size_t J = T.numJoints;
    stageTable(T.parent, J, cursor); stageTable(T.offset, 3 * J, cursor); stageTable(T.prerot, 4 * J, cursor);
    stageTable(T.ptOuter, 7 * J + 1, cursor); stageTable(T.ptInner, T.ptNnz, cursor); stageTable(T.ptVals, T.ptNnz, cursor);
    stageTable(T.ptOffsets, 7 * J, cursor);
    stageTable(T.levelStart, T.numLevels + 1, cursor); stageTable(T.levelJoints, J, cursor);
    stageTable(T.efs, T.numEf, cursor); stageTable(T.units, T.numUnits, cursor); stageTable(T.limitData, T.numLimitData, cursor);
    if constexpr (kStage == 1) { stageTable(T.cells, T.numCells, cursor); stageTable(T.contribs, T.numContribs, cursor); }
    __syncthreads();
  }

  for (int b = blockIdx.x * groupsPerCta + group; b < a.batch; b += gridDim.x * groupsPerCta) {
    if (a.active != nullptr && a.active[b] == 0) continue; // (the whole group skips together)
    const float* theta = a.theta + size_t(b) * a.ldTheta;
    for (int i = gl; i < T.numParams; i += gs) th[i] = theta[i];
    groupSync();
    if (W == 1) { // touch the next instance's parameters now: by the time this one is done they sit in L1 / L2 (one warp = one 880-byte row)
      const int bn = b + gridDim.x * groupsPerCta;
      if (bn < a.batch) { const float* tn = a.theta + size_t(bn) * a.ldTheta; for (int i = gl * 8; i < T.numParams; i += gs * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(tn + i)); }
    }
    // SkeletonState::set in three data-parallel passes (ik_device.cuh): every joint's local part at once (its seven joint parameters
    // come straight from theta: ParameterTransform::apply row by row), ~50 dependent flops per tree level, then the derivative axes
    // (one warp per instance: three rounds of lanes = joints, each walking its seven rows; several warps per instance: lanes = rows first)
    if constexpr (W > 1) {
      for (int row = gl; row < T.numJoints * kParametersPerJoint; row += gs) jp[row] = jointParameterRow(T, row, th);
      groupSync();
      for (int j = gl; j < T.numJoints; j += gs) fkLocal<kJacobian>(T, j, jp, js);
    } else {
      for (int j = gl; j < T.numJoints; j += gs) fkLocalFromTheta<kJacobian>(T, j, th, js);
    }
    groupSync();
    for (int lvl = 1; lvl < T.numLevels; ++lvl) { // level 0 = roots: world = local
      const int end = T.levelStart[lvl + 1];
      for (int k = T.levelStart[lvl] + gl; k < end; k += gs) fkCompose(T, T.levelJoints[k], js);
      groupSync();
    }
    if (kJacobian) {
      for (int i = gl; i < 3 * T.numJoints; i += gs) fkAxis(T, i / 3, i % 3, js);
      groupSync();
    }
    if (a.stateOut != nullptr) {
      float* so = a.stateOut + size_t(b) * T.numJoints * 8;
      for (int i = gl; i < T.numJoints * 8; i += gs) so[i] = js[(i >> 3) * kJointStateStride + (i & 7)];
    }
    const float* targets = a.targets + size_t(b) * T.targetStride;
    const float* cw = a.cweights + (T.weightsPerInstance ? size_t(b) * T.numWeights : 0);
    float* J = kJacobian ? a.jacobian + size_t(b) * T.jacobianStride : nullptr;
    // the residual is the last column of the device matrix, or follows the strips
    float* residual = kJacobian ? J + (T.stripMode ? size_t(T.residOff) : size_t(T.numCols) * T.ldJ) : nullptr;
    double err = 0.0;
    for (int u = gl; u < T.numUnits; u += gs) err += (double)evalUnit<kJacobian>(T, u, th, jp, js, targets, cw, rec, residual);
    err = warpSum(err);
    if (W > 1 && lane == 0) errSlots[warp % W] = err;
    groupSync();
    if (kJacobian)
      for (int c = gl; c < T.numCells; c += gs) jacobianCell(T, c, js, rec, targets, J);
    if (gl == 0) {
      if (W > 1) { err = 0.0; for (int w = 0; w < W; ++w) err += errSlots[w]; } // fixed order: deterministic
      // getError() rounds through float (skeleton_solver_function.cpp:82); the Jacobian pass keeps double
      a.errors[b] = kJacobian ? err : (double)(float)err;
    }
    groupSync();
  }
}

static int g_numSms = 0;
static int g_maxSmemOptin = 0;
static int g_maxSmemPerSm = 0;

cudaError_t launchSweep(const SweepArgs& a0, bool jacobian, cudaStream_t stream) {
  SweepArgs a = a0;
  const size_t budget = size_t(g_maxSmemOptin);
  // Persistent CTAs. As many instances in flight as shared memory holds next to the staged tables: all of them when they fit with at
  // least two instances, else all but the cell / contribution records (large rigs), else none. One warp per instance, up to
  // kSweepMaxWarps; when fewer than 16 instances fit, several warps share one instance (up to kSweepMaxWarps warps per CTA).
  const size_t per1 = sweepSmemPerInstance(a.T, 8);
  a.stageTables = (2 * per1 + sweepTableBytes(a.T, 1) + 16 <= budget) ? 1 : (2 * per1 + sweepTableBytes(a.T, 2) + 16 <= budget) ? 2 : 0;
  const size_t tableBytes = sweepTableBytes(a.T, a.stageTables) + 16;
  int groups = int((budget - tableBytes) / per1);
  if (groups < 1) return cudaErrorInvalidConfiguration;
  const int groupsOneWarp = int((budget - tableBytes) / sweepSmemPerInstance(a.T, 1)); // (no joint-parameter array)
  int W = 1;
  if (groupsOneWarp >= 16) {
    // The instances of an SM are dealt to its warps round by round, and from about a dozen warps on the kernel is throughput bound
    // (measured on the cfg3 shard: 0.049 / 0.059 / 0.095 ms per round with 16 / 19 / 28 warps): the time is rounds x warps, i.e. the
    // padded instance count. Pick the warp count that wastes the fewest slots (8192 instances on 148 SMs: 19 warps x 3 rounds = 8436
    // slots; 4096 instances: 14 warps x 2 rounds = 4144); a 28-warp variant (73 registers) spilled and lost more than its two rounds won.
    const int sms = std::max(g_numSms, 1);
    const int maxG = std::min(groupsOneWarp, kSweepMaxWarps), minG = std::min(12, maxG);
    long best = -1;
    if (a.batch < sms * minG) { groups = std::max(1, (a.batch + sms - 1) / sms); best = 0; } // a small batch: spread it over the SMs
    for (int gcand = maxG; gcand >= minG && best != 0; --gcand) {
      const long rounds = (a.batch + long(sms) * gcand - 1) / (long(sms) * gcand);
      if (best < 0 || rounds * gcand < best) { best = rounds * gcand; groups = gcand; }
    }
  } else {
    if (groups > 15) groups = 15; // named barriers 1..15
    while (W < 8 && groups * W * 2 <= kSweepMaxWarps) W *= 2; // these instances wait on L1 / L2 and on each other's barriers: as many warps as fit
  }
  a.warpsPerInstance = W;
  const int warps = groups * W;
  const size_t per = sweepSmemPerInstance(a.T, W);
  const size_t smem = per * groups + tableBytes;
  if (smem > budget) return cudaErrorInvalidConfiguration;
  const int ctasNeeded = (a.batch + groups - 1) / groups;
  int ctasPerSm = (int)((size_t(g_maxSmemPerSm)) / (smem + 1024));
  if (ctasPerSm < 1) ctasPerSm = 1;
  if (ctasPerSm * warps > 32) ctasPerSm = 32 / warps > 0 ? 32 / warps : 1;
  int grid = g_numSms * ctasPerSm;
  if (grid > ctasNeeded) grid = ctasNeeded;
  if (grid < 1) grid = 1;
  auto launch = [&](auto kernel) -> cudaError_t {
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    kernel<<<grid, warps * 32, smem, stream>>>(a);
    return cudaGetLastError();
  };
  auto pick = [&](auto jac, auto staged) -> cudaError_t {
    constexpr bool kJ = decltype(jac)::value;
    constexpr int kS = decltype(staged)::value;
    switch (W) {
      case 1: return launch(sweepKernel<kJ, 1, kS>);
      case 2: return launch(sweepKernel<kJ, 2, kS>);
      case 4: return launch(sweepKernel<kJ, 4, kS>);
      default: return launch(sweepKernel<kJ, 8, kS>);
    }
  };
  auto pickStage = [&](auto jac) -> cudaError_t {
    switch (a.stageTables) {
      case 1: return pick(jac, std::integral_constant<int, 1>{});
      case 2: return pick(jac, std::integral_constant<int, 2>{});
      default: return pick(jac, std::integral_constant<int, 0>{});
    }
  };
  return jacobian ? pickStage(std::true_type{}) : pickStage(std::false_type{});
}


// ------------------------------------------------------------------------------------------------
// K2 (SIMT validation path): H[i][j] = sum_k J[k][cols[i]] J[k][cols[j]] for i >= j; g = J^T r.
// grid = (lower-triangular 64x64 tile pairs, B); 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------------
constexpr int kJtjTile = 64;
constexpr int kJtjKc = 32;

__global__ void __launch_bounds__(256) jtjSimtKernel(const JtJArgs a) {
  const int b = blockIdx.y;
  if (a.active != nullptr && a.active[b] == 0) return;
  // decode lower-triangular tile index
  int t = blockIdx.x, ti = 0;
  while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
  const int tj = t - ti * (ti + 1) / 2;
  __shared__ float As[kJtjTile][kJtjKc + 1];
  __shared__ float Bs[kJtjTile][kJtjKc + 1];
  __shared__ float rs[kJtjKc];
  const float* J = a.jacobian + size_t(b) * (a.numCols + 1) * a.ldJ;
  const float* r = J + size_t(a.numCols) * a.ldJ;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float gacc = 0.f;
  const bool diag = (ti == tj);
  for (int k0 = 0; k0 < a.kRows; k0 += kJtjKc) {
    // load 64 columns x 32 rows for both operands (lane -> consecutive k: coalesced 128B per column)
    for (int idx = threadIdx.x; idx < kJtjTile * kJtjKc; idx += 256) {
      const int c = idx / kJtjKc, kk = idx % kJtjKc;
      const int ia = ti * kJtjTile + c, ib = tj * kJtjTile + c;
      const int k = k0 + kk;
      As[c][kk] = (ia < a.ns && k < a.kRows) ? J[size_t(ia) * a.ldJ + k] : 0.f;
      Bs[c][kk] = (ib < a.ns && k < a.kRows) ? J[size_t(ib) * a.ldJ + k] : 0.f;
    }
    if (threadIdx.x < kJtjKc) rs[threadIdx.x] = (k0 + threadIdx.x < a.kRows) ? r[k0 + threadIdx.x] : 0.f;
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kJtjKc; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[ty * 4 + i][kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[tx * 4 + j][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (diag && threadIdx.x < kJtjTile)
      for (int kk = 0; kk < kJtjKc; ++kk) gacc = fmaf(As[threadIdx.x][kk], rs[kk], gacc);
    __syncthreads();
  }
  float* H = a.H + size_t(b) * a.hStride;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = ti * kJtjTile + ty * 4 + i, gj = tj * kJtjTile + tx * 4 + j;
      if (gi < a.ns && gj <= gi) { H[size_t(gj) * a.ldH + gi] = acc[i][j]; H[size_t(gi) * a.ldH + gj] = acc[i][j]; }
    }
  if (diag && threadIdx.x < kJtjTile) {
    const int gi = ti * kJtjTile + threadIdx.x;
    if (gi < a.ns) {
      H[size_t(gi) * a.ldH + a.ns] = gacc; H[size_t(a.ns) * a.ldH + gi] = gacc;
      if (a.g != nullptr) a.g[size_t(b) * a.ldG + gi] = gacc;
    }
  }
}

cudaError_t launchJtJSimt(const JtJArgs& a, cudaStream_t stream) {
  const int tiles = (a.ns + kJtjTile - 1) / kJtjTile;
  dim3 grid(tiles * (tiles + 1) / 2, a.batch);
  jtjSimtKernel<<<grid, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K3: one CTA per instance. Matrix A = [H + lambda I ; g^T] ((ns+1) x lda, row-major, lower part)
// is factored in place with Eigen's LLT structure (ik_chol.cuh); the appended row turns into
// y = L^-1 g for free; back substitution gives delta.
// ------------------------------------------------------------------------------------------------
// Common tail of both Cholesky kernels: delta, parameter update (skeleton_solver_function.cpp:153-159),
// g.delta for the subset line search, status and the SolverT bookkeeping (solver.cpp:92-122).
// dsub / gsub are indexed by subset position. Every thread of the CTA must call it.
// dsub: the step per device column, or (slotOf != nullptr) per elimination slot with slotOf[i] = slot of device column i
__device__ void cholFinish(const CholArgs& a, int b, int n, const float* dsub, const float* gsub, bool failed, const int32_t* slotOf = nullptr) {
  const int tid = threadIdx.x;
  float* theta = a.theta + size_t(b) * a.ldTheta;
  float part = 0.f;
  for (int i = tid; i < n; i += blockDim.x) {
    const int c = a.cols[i];
    const float d = c >= 0 ? dsub[slotOf != nullptr ? slotOf[i] : i] : 0.f; // c < 0: all-zero alignment column of the scheduled layout (ik_chol_sched.h)
    a.delta[size_t(b) * n + i] = d;
    if (c >= 0) part += gsub[i] * d;
    if (a.applyUpdate && c >= 0) theta[c] -= d;
  }
  if (a.gradDotDelta != nullptr) {
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int w = 0; w < int(blockDim.x >> 5); ++w) s += red[w];
      a.gradDotDelta[b] = s;
    }
  }
  if (tid == 0) {
    if (failed && a.status[b] == 0) a.status[b] = 1; // MB2_INSTANCE_CHOLESKY_BREAKDOWN
    if (a.bookkeeping) {
      const double error = a.errors[b], last = a.lastErrors[b];
      if (a.history != nullptr) a.history[size_t(b) * a.maxIterations + a.iteration] = error;
      const bool converged = fabs(last - error) / (fabs(error) + (double)FLT_MIN) <= (double)(a.threshold * FLT_EPSILON);
      a.iterations[b] = a.iteration + 1;
      if ((a.iteration >= a.minIterations && converged) || a.iteration + 1 >= a.maxIterations) a.active[b] = 0;
      else { a.lastErrors[b] = error; atomicAdd(a.activeCount, 1); }
    }
  }
}

template <int NB>
__global__ void __launch_bounds__(kCholThreads) choleskyKernel(const CholArgs a, const int useSmemMatrix) {
  extern __shared__ float smem[];
  const int b = blockIdx.x;
  if (a.active[b] == 0) return;
  const int n = a.ns;
  const int tid = threadIdx.x;
  float* Hg = a.H + size_t(b) * a.hStride;
  CholCtx ctx;
  ctx.n = n;
  int ldp = ((n + 1 + 3) & ~3) + 4;
  float* P = smem;                       // [NB][ldp] transposed panel
  float* gsave = P + NB * ldp;           // [n] copy of Jtr (for g.delta)
  int* flags = reinterpret_cast<int*>(gsave + ((n + 3) & ~3)); // [0] = fail index+1
  float* As = reinterpret_cast<float*>(flags + 4);
  if (useSmemMatrix) {
    ctx.lda = n | 1;
    ctx.A = As;
    // source is column-major (coalesced along i); smem target is row-major with odd stride
    for (int idx = tid; idx < (n + 1) * n; idx += kCholThreads) {
      const int j = idx / (n + 1), i = idx - j * (n + 1);
      if (i >= j) {
        float v = Hg[size_t(j) * a.ldH + i];
        if (i == j) v += a.regularization; // gauss_newton_solver.cpp:248
        As[i * ctx.lda + j] = v;
      }
    }
  } else {
    // matrix too large for shared memory: factor in place in global memory (L2-resident). K2 wrote the lower
    // triangle column-major (element (i,j) at [j*ldH + i]); mirror it so that A(i,j) = Hg[i*ldH + j] reads row-major.
    ctx.lda = a.ldH;
    ctx.A = Hg;
    for (int idx = tid; idx < (n + 1) * n; idx += kCholThreads) {
      const int j = idx / (n + 1), i = idx - j * (n + 1);
      if (i > j) Hg[size_t(i) * a.ldH + j] = Hg[size_t(j) * a.ldH + i];
    }
    for (int i = tid; i < n; i += kCholThreads) Hg[size_t(i) * a.ldH + i] += a.regularization;
  }
  ctx.P = P;
  ctx.ldp = ldp;
  ctx.fail = flags;
  if (tid == 0) flags[0] = 0;
  for (int i = tid; i < n; i += kCholThreads) gsave[i] = Hg[size_t(i) * a.ldH + n];
  __syncthreads();

  const int blockSize = cholBlockSize(n, NB);
  for (int k = 0; k < n; k += blockSize) {
    const int bs = min(blockSize, n - k);
    if (tid < 32) { // unblocked LLT of the diagonal block by warp 0 (left-looking, Eigen llt_inplace::unblocked)
      for (int jj = 0; jj < bs; ++jj) {
        const float x = cholDiagPivot(ctx, k, jj);
        __syncwarp();
        if (!(x > 0.f)) { if (tid == 0) flags[0] = k + jj + 1; break; }
        cholDiagColumn(ctx, k, bs, jj, x, tid);
        __syncwarp();
      }
    }
    __syncthreads();
    if (flags[0] != 0) break; // Eigen returns early: the rest of the matrix stays as it is
    cholPanelSolve<NB>(ctx, k, bs, tid, kCholThreads);
    __syncthreads();
    cholTrailingUpdate<NB>(ctx, k, bs, tid, kCholThreads);
    __syncthreads();
  }
  float* y = ctx.A + size_t(n) * ctx.lda; // appended row
  if (flags[0] != 0) {
    // finish the forward substitution with whatever the lower triangle holds (LLT::solve after a failed compute)
    if (tid == 0) cholForwardFrom(ctx, cholCompletedColumns(flags[0] - 1, blockSize), y);
    __syncthreads();
  }
  // back substitution L^T x = y, 32 columns at a time from the bottom
  for (int i0 = ((n - 1) / 32) * 32; i0 >= 0; i0 -= 32) {
    const int nb = min(32, n - i0);
    if (tid < 32) {
      float yv = tid < nb ? y[i0 + tid] : 0.f;
      float dinv = tid < nb ? 1.f / ctx.A[size_t(i0 + tid) * ctx.lda + i0 + tid] : 0.f;
      for (int kk = nb - 1; kk >= 0; --kk) {
        const float xk = __shfl_sync(0xffffffffu, yv, kk) * __shfl_sync(0xffffffffu, dinv, kk);
        if (tid == kk) yv = xk;
        else if (tid < kk) yv -= ctx.A[size_t(i0 + kk) * ctx.lda + i0 + tid] * xk;
      }
      if (tid < nb) y[i0 + tid] = yv;
    }
    __syncthreads();
    for (int i = tid; i < i0; i += kCholThreads) {
      float s = y[i];
      for (int kk = 0; kk < nb; ++kk) s -= ctx.A[size_t(i0 + kk) * ctx.lda + i] * y[i0 + kk];
      y[i] = s;
    }
    __syncthreads();
  }
  cholFinish(a, b, n, y, gsave, flags[0] != 0);
}

static size_t cholSmemBytes(int n, int NB, bool matrixInSmem) {
  const size_t ldp = ((n + 1 + 3) & ~3) + 4;
  size_t s = sizeof(float) * (NB * ldp + ((n + 3) & ~3)) + 4 * sizeof(int);
  if (matrixInSmem) s += sizeof(float) * size_t(n + 1) * (n | 1);
  return s;
}

cudaError_t launchCholesky(const CholArgs& a, cudaStream_t stream) {
  const int n = a.ns;
  const int eig = cholBlockSize(n, 32);
  const int NB = eig <= 8 ? 8 : (eig <= 16 ? 16 : 32);
  bool inSmem = cholSmemBytes(n, NB, true) <= size_t(g_maxSmemOptin);
  const size_t smem = cholSmemBytes(n, NB, inSmem);
  cudaError_t e = cudaSuccess;
#define MB2_LAUNCH_CHOL(NBV)                                                                                   \
  e = cudaFuncSetAttribute(choleskyKernel<NBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));       \
  if (e != cudaSuccess) return e;                                                                              \
  choleskyKernel<NBV><<<a.batch, kCholThreads, smem, stream>>>(a, inSmem ? 1 : 0);
  if (NB == 8) { MB2_LAUNCH_CHOL(8) } else if (NB == 16) { MB2_LAUNCH_CHOL(16) } else { MB2_LAUNCH_CHOL(32) }
#undef MB2_LAUNCH_CHOL
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K2s: tile-sparse Gram. A skeleton Jacobian is mostly structural zeros (a row only touches the parameters of one
// root-to-constraint chain), so the stored 16x16 tiles of J^T J are accumulated from the non-zero 4-row x 16-column strips only
// (GramPlan): ~20x fewer multiply-adds than the dense product, fp32-class accuracy (three-term TF32 split on mma.sync, the lo*lo term dropped: ~2^-21 relative). One CTA per instance: every strip arrives as one
// TMA box of the K-major Jacobian, a warp owns a tile at a time, the output is already in the Cholesky kernel's tile layout.
// ------------------------------------------------------------------------------------------------
constexpr int kGramThreads = 32 * kGramWarps;

size_t gramTilesSmemBytes(size_t stripStride, int blobInts) { return 128 + sizeof(float) * (stripStride + 64) + 16 + sizeof(int32_t) * size_t((blobInts + 3) & ~3); }

// kThreads = 256 (eight warps, the width the tile-order table is dealt for), or 512 when the strips of an instance leave room for only
// one CTA per SM anyway (bodyhands300: 155 KB): the sixteen warps then walk the same table two rounds at a time.
template <int kThreads>
__global__ void __launch_bounds__(kThreads) gramTilesKernel(const GramArgs a) {
  extern __shared__ __align__(16) float gramSmem[];
  const int b = blockIdx.x;
  if (a.active != nullptr && a.active[b] == 0) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, hw = tid >> 4, hl = tid & 15;
  float* strips = gramSmem + (((128u - (smemAddr(gramSmem) & 127u)) & 127u) >> 2);
  float* resid = strips + a.residOff;
  // smem: [strips | residual] as in global memory, then the all-zero strip that pads odd pair lists (index zeroStrip), barrier, tables
  float* zero = strips + a.stripStride;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(zero + 64);
  const uint32_t barAddr = smemAddr(bar);
  const uint32_t total = uint32_t(a.stripStride) * 4u;
  if (tid < 64) zero[tid] = 0.f;
  if (tid == 0) {
    mbarInit(barAddr, 1);
    fenceBarrierInit();
    mbarExpectTx(barAddr, total);
    const char* src = reinterpret_cast<const char*>(a.strips + size_t(b) * a.stripStride);
    for (uint32_t off = 0; off < total; off += 16384u) bulkLoad(smemAddr(strips) + off, src + off, total - off < 16384u ? total - off : 16384u, barAddr);
  }
  // the plan tables are read in dependent chains (tile -> pair range -> strips): stage them in shared memory while the copy is in flight
  int32_t* tab = reinterpret_cast<int32_t*>(bar + 2);
  for (int i = tid; i < a.blobInts; i += kThreads) tab[i] = __ldg(a.blob + i);
  __syncthreads();
  mbarWaitRelaxed(barAddr, 0);
  const int32_t* tileOrder = tab + a.offTileOrder, *tileQuadStart = tab + a.offTilePairStart, *quads = tab + a.offPairA; // (blob tables are 16-byte aligned)
  const int32_t* colStripStart = tab + a.offColStripStart, *colStrip = tab + a.offColStrip, *stripRow = tab + a.offStripRow, *tileInfo = tab + a.offTileInfo;
  float* out = a.out + size_t(b) * a.outStride;
  for (int ti = warp; ti < a.numOrder; ti += kThreads / 32) { // ([rounds][8] table: warp w + 8 takes the odd rounds of warp w's list)
    const int t = tileOrder[ti];
    if (t < 0) continue;
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    gramTileAccumulate(strips, quads, tileQuadStart[t], tileQuadStart[t + 1], lane, acc);
    gramTileStore(out + size_t(t) * 256, acc, tileInfo[t], a.regularization, lane);
  }
  float* y = out + size_t(a.numTiles) * 256;
  for (int K = hw; K < a.numTileCols; K += kThreads / 16)
    y[16 * K + hl] = gramVectorEntry(strips, resid, colStrip, stripRow, colStripStart[K], colStripStart[K + 1], hl);
}

cudaError_t launchGramTiles(const GramArgs& a, cudaStream_t stream) {
  const size_t smem = gramTilesSmemBytes(a.stripStride, a.blobInts);
  if (smem > size_t(g_maxSmemOptin) || (a.stripStride & 3) != 0) return cudaErrorInvalidConfiguration;
  const bool wide = 2 * (smem + 1024) > size_t(g_maxSmemPerSm); // one CTA per SM anyway: give it 16 warps
  cudaError_t e = wide ? cudaFuncSetAttribute(gramTilesKernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))
                       : cudaFuncSetAttribute(gramTilesKernel<kGramThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  if (wide) gramTilesKernel<512><<<a.batch, 512, smem, stream>>>(a);
  else gramTilesKernel<kGramThreads><<<a.batch, kGramThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K3 (scheduled): level-scheduled tile-sparse Cholesky of the permuted system (ik_chol_sched.h).
// One CTA per instance; tiles, right-hand side and scratch live in shared memory (<= ~64 KB for the
// humanoid rig => several CTAs per SM hide each other's latencies).
// ------------------------------------------------------------------------------------------------
// 256 threads and three CTAs per SM when the tiles of one instance take a third of shared memory (humanoid-size rigs); 512 threads in
// the single resident CTA when one instance needs more than half of it (body + hands)

size_t choleskyScheduledSmemBytes(int n, int nPad, int numTiles, int blobInts) {
  return 1024 /*tile storage is aligned to the TMA swizzle atom*/ + sizeof(float) * (size_t(numTiles) * 256 + size_t(nPad) + 2 * size_t((n + 3) & ~3)) +
         sizeof(int32_t) * size_t((blobInts + 3) & ~3) + 32;
}

// kProfile: per-phase cycle counters of block 0 (MB2_CHOL_PROFILE=1); a separate instantiation so that the production kernel does not
// carry the counters in its register budget
template <int kSchedThreads, bool kProfile>
__global__ void __launch_bounds__(kSchedThreads, kSchedThreads == 256 ? 3 : 1) choleskyScheduledKernel(const __grid_constant__ CUtensorMap hmap, const CholArgs a, const CholSchedDev Sg) {
  extern __shared__ __align__(16) float smemRaw[];
  const int b = blockIdx.x;
  if (a.active[b] == 0) return;
  const int n = a.ns, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, hw = tid >> 4, hl = tid & 15;
  const unsigned hmask = 0xFFFFu << (16 * ((tid >> 4) & 1));
  float* tiles = smemRaw + (((1024u - (smemAddr(smemRaw) & 1023u)) & 1023u) >> 2); // 1 KB aligned in the shared window (TMA swizzle atom)
  float* y = tiles + size_t(Sg.numTiles) * 256;
  float* gsub = y + Sg.nPad;
  float* dsub = gsub + ((n + 3) & ~3);
  int32_t* blob = reinterpret_cast<int32_t*>(dsub + ((n + 3) & ~3));
  int* flags = reinterpret_cast<int*>(blob + ((Sg.blobInts + 3) & ~3));
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(flags + 2);
  long long pc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, pt = clock64();
#define MB2_PROF(k) if constexpr (kProfile) { const long long now = clock64(); pc[k] += now - pt; pt = now; }
  // Prologue: everything this instance reads arrives asynchronously on one mbarrier -- the schedule tables and J^T r as 1-D
  // bulk copies, every stored tile as one TMA box (16 rows x 64 bytes of the row-major upper triangle of H, written
  // as swizzled 64-byte rows: SWIZZLE_64B, see tmaBoxIdx; cholConvertBox then rewrites each box in the fragment layout).
  const uint32_t barAddr = smemAddr(bar);
  const uint32_t blobBytes = uint32_t((Sg.blobInts + 3) & ~3) * 4u, gBytes = uint32_t(a.ldG) * 4u;
  const bool fromGram = a.tilesIn != nullptr; // tiles (+ lambda, identity extension) and the slot-ordered J^T r come ready-made from the Gram kernel
  const uint32_t tileBytes = uint32_t(Sg.numTiles) * 1024u, yBytes = uint32_t(Sg.nPad) * 4u;
  if (tid == 0) {
    flags[0] = 0;
    mbarInit(barAddr, 1);
    fenceBarrierInit();
    mbarExpectTx(barAddr, blobBytes + tileBytes + (fromGram ? yBytes : gBytes));
  }
  __syncthreads();
  if (fromGram) {
    if (tid == 0) {
      bulkLoad(smemAddr(blob), Sg.blob, blobBytes, barAddr);
      const float* src = a.tilesIn + size_t(b) * a.tilesStride;
      const uint32_t total = tileBytes + yBytes; // y follows the tiles in both layouts
      for (uint32_t off = 0; off < total; off += 16384u)
        bulkLoad(smemAddr(tiles) + off, reinterpret_cast<const char*>(src) + off, total - off < 16384u ? total - off : 16384u, barAddr);
    }
  } else {
    if (tid == 0) {
      bulkLoad(smemAddr(blob), Sg.blob, blobBytes, barAddr);
      bulkLoad(smemAddr(gsub), a.g + size_t(b) * a.ldG, gBytes, barAddr);
    }
    for (int t = tid; t < Sg.numTiles; t += kSchedThreads) { // one thread per tile: the table reads overlap, the compiler serialises the TMA issue per warp
      const int gi0 = __ldg(Sg.tileInfo + 3 * t), gj0 = __ldg(Sg.tileInfo + 3 * t + 1);
      tmaLoad3d(smemAddr(tiles + size_t(t) * 256), &hmap, gi0, gj0, b, barAddr);
    }
  }
  MB2_PROF(6)
  mbarWaitRelaxed(barAddr, 0); // 256 spinning threads would take issue slots from the other CTAs of the SM
  MB2_PROF(7)
  const CholSchedDev S = rebaseSchedule(Sg, blob);
  if (fromGram) {
    for (int s = tid; s < S.nPad; s += kSchedThreads) {
      const int p = S.perm[s];
      if (p >= 0) gsub[p] = y[s];
    }
  } else {
    // the boxes landed as swizzled 64-byte rows: a warp per tile turns them into the fragment layout (identity extension on padding)
    for (int t = warp; t < S.numTiles; t += kSchedThreads / 32) cholConvertBox(tiles + size_t(t) * 256, S.tileInfo[3 * t + 2], lane);
    for (int s = tid; s < S.nPad; s += kSchedThreads) {
      const int p = S.perm[s];
      y[s] = p >= 0 ? gsub[p] : 0.f;
    }
  }
  __syncthreads();
  MB2_PROF(8)
  if (!fromGram)
    for (int s = tid; s < S.nPad; s += kSchedThreads)
      if (S.perm[s] >= 0) tiles[size_t(S.diagTile[s >> 4]) * 256 + tileIdx(s & 15, s & 15)] += a.regularization; // gauss_newton_solver.cpp:248
  __syncthreads();
  MB2_PROF(0)

  for (int L = 0; L < S.numLevels; ++L) {
    // A: diagonal tiles of this level (one warp each) + forward solve of their rhs block
    for (int ci = S.levelColStart[L] + warp; ci < S.levelColStart[L + 1]; ci += kSchedThreads / 32) {
      const int K = S.levelCols[ci];
      cholDiagTile(tiles + size_t(S.diagTile[K]) * 256, y + 16 * K, lane, a.regularization, flags);
    }
    __syncthreads();
    MB2_PROF(1)
    // B: panel tiles
    for (int pi = S.levelPanelStart[L] + warp; pi < S.levelPanelStart[L + 1]; pi += kSchedThreads / 32) {
      float* ptile = tiles + size_t(S.panelTile[pi]) * 256;
      float x[2][4];
      cholPanelProduct(ptile, tiles + size_t(S.panelDiag[pi]) * 256, lane, x);
      cholPanelStore(ptile, lane, x);
    }
    __syncthreads();
    MB2_PROF(2)
    // C: update tasks (warp each) and rhs updates (half-warp each)
    {
      const int32_t* oStart = kSchedThreads == 256 ? S.levelOrderStart8 : S.levelOrderStart16, *order = kSchedThreads == 256 ? S.taskOrder8 : S.taskOrder16;
      for (int oi = oStart[L] + warp; oi < oStart[L + 1]; oi += kSchedThreads / 32) { const int ti = order[oi]; if (ti >= 0) cholUpdateTask(tiles, S, ti, lane); }
    }
    for (int vi = S.levelVTaskStart[L] + hw; vi < S.levelVTaskStart[L + 1]; vi += kSchedThreads / 16) cholVectorTask(tiles, y, S, vi, hl);
    __syncthreads();
    MB2_PROF(3)
  }
  for (int L = S.numLevels - 1; L >= 0; --L) {
    for (int ci = S.levelColStart[L] + warp; ci < S.levelColStart[L + 1]; ci += kSchedThreads / 32) cholBackwardColumn(tiles, y, S, S.levelCols[ci], lane);
    __syncthreads();
  }
  MB2_PROF(4)
  for (int i = tid; i < S.nPad; i += kSchedThreads) { const int p = S.perm[i]; if (p >= 0) dsub[p] = y[i]; }
  __syncthreads();
  cholFinish(a, b, n, dsub, gsub, flags[0] != 0);
  MB2_PROF(5)
  if (kProfile && b == 0 && tid == 0)
    printf("chol-profile (cycles, block 0): blob %lld issue %lld wait %lld lambda %lld diag %lld panel %lld update %lld backward %lld finish %lld | levels %d tiles %d\n", pc[6], pc[7],
           pc[8], pc[0], pc[1], pc[2], pc[3], pc[4], pc[5], S.numLevels, S.numTiles);
#undef MB2_PROF
}

cudaError_t launchCholeskyScheduled(const CholArgs& a, const CholSchedDev& sched, cudaStream_t stream) {
  const size_t smem = choleskyScheduledSmemBytes(a.ns, sched.nPad, sched.numTiles, sched.blobInts);
  if (smem > size_t(g_maxSmemOptin)) return cudaErrorInvalidConfiguration;
  const bool wide = 2 * (smem + 1024) > size_t(g_maxSmemPerSm); // one CTA per SM anyway: give it 16 warps
  const bool prof = (a.profile & 1) != 0;
  cudaError_t e;
  if (wide) e = prof ? cudaFuncSetAttribute(choleskyScheduledKernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))
                     : cudaFuncSetAttribute(choleskyScheduledKernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  else e = prof ? cudaFuncSetAttribute(choleskyScheduledKernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))
                : cudaFuncSetAttribute(choleskyScheduledKernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  if (a.tilesIn == nullptr && (a.g == nullptr || a.ldG != cholGradientLd(a.ns))) return cudaErrorInvalidValue;
  if ((a.ldH & 3) != 0 || (a.hStride & 3) != 0 || (a.tilesStride & 3) != 0) return cudaErrorInvalidValue;
  CUtensorMap hmap; // H as [batch][ns + 1][ldH]; one 16 x 16 box per stored tile (unused, but still a valid map, when the tiles come from the Gram kernel)
  const bool fromGram = a.tilesIn != nullptr;
  const uint64_t dims[3] = {uint64_t(fromGram ? 256 : a.ldH), uint64_t(fromGram ? sched.numTiles : a.ns + 1), uint64_t(a.batch)};
  const uint64_t strides[2] = {uint64_t(fromGram ? 256 : a.ldH) * sizeof(float), uint64_t(fromGram ? a.tilesStride : a.hStride) * sizeof(float)};
  const uint32_t box[3] = {16u, 16u, 1u};
  e = makeTensorMap3d(&hmap, fromGram ? a.tilesIn : a.H, dims, strides, box, 64);
  if (e != cudaSuccess) return e;
  if (wide && prof) choleskyScheduledKernel<512, true><<<a.batch, 512, smem, stream>>>(hmap, a, sched);
  else if (wide) choleskyScheduledKernel<512, false><<<a.batch, 512, smem, stream>>>(hmap, a, sched);
  else if (prof) choleskyScheduledKernel<256, true><<<a.batch, 256, smem, stream>>>(hmap, a, sched);
  else choleskyScheduledKernel<256, false><<<a.batch, 256, smem, stream>>>(hmap, a, sched);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K2s + K3 fused (default on the tile path): one CTA per instance, three CTAs per SM.
//   prologue   strips + residual of the instance, the Gram plan and the Cholesky schedule arrive as bulk copies on one mbarrier
//   Gram       a warp owns a tile at a time (longest-first assignment); finished accumulators are parked in TMEM
//              (tcgen05.st) because the tiles are about to overwrite the strips they are computed from
//   restore    tcgen05.ld -> 16x16 tiles of J^T J + lambda I in the Cholesky layout, over the dead strips
//   Cholesky   level-scheduled tile factorisation, both substitutions, theta -= delta, SolverT bookkeeping (cholFinish)
// The stored tiles (0.45 GB per iteration on the cfg3 shard) never exist in HBM and the second launch is gone.
// ------------------------------------------------------------------------------------------------
size_t gramCholeskySmemBytes(size_t stripStride, int gramBlobInts, int n, int nPad, int numTiles, int schedBlobInts) {
  (void)gramBlobInts; (void)schedBlobInts; // the plan tables stay in global memory (L1): see the kernel
  const size_t uni = std::max<size_t>(size_t(numTiles) * 256, stripStride + 64);
  return 128 + sizeof(float) * (((uni + 3) & ~size_t(3)) + size_t(nPad) + size_t((n + 3) & ~3)) + 32;
}

// Shared memory holds only what is per instance: the strip / tile union, the right-hand side in slot order and J^T r per device column
// (cfg3: 56.4 KB), so that FOUR CTAs fit on an SM (57 KB each, 64 registers per thread); the Gram plan and the Cholesky schedule
// (10 KB, identical for every CTA) are read from global memory through L1 by the same device functions. The per-instance critical
// path is a chain of dependent phases no wider than a few warps: instances in flight per SM are what buys throughput.
template <bool kProfile>
__global__ void __launch_bounds__(kGramThreads, 4) gramCholeskyKernel(const GramCholArgs a, const CholSchedDev S, const int tmemColumns) {
  extern __shared__ __align__(16) float gcSmem[];
  const GramArgs& g = a.g;
  const CholArgs& c = a.c;
  const int b = blockIdx.x;
  if (c.active[b] == 0) return;
  const int n = c.ns, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, hw = tid >> 4, hl = tid & 15;
  const unsigned hmask = 0xFFFFu << (16 * ((tid >> 4) & 1));
  float* U = gcSmem + (((128u - (smemAddr(gcSmem) & 127u)) & 127u) >> 2); // strips | residual | zero strip, later the tiles
  const size_t tileFloats = size_t(S.numTiles) * 256, sweepFloats = g.stripStride + 64;
  const size_t uni = ((tileFloats > sweepFloats ? tileFloats : sweepFloats) + 3) & ~size_t(3);
  float* strips = U;
  float* resid = U + g.residOff;
  float* tiles = U;
  float* y = U + uni;
  float* gsub = y + S.nPad;
  int* flags = reinterpret_cast<int*>(gsub + ((n + 3) & ~3));
  uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(flags + 2);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(flags + 4);
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = 0;
  if constexpr (kProfile) pt = clock64();
#define MB2_GC(k) if constexpr (kProfile) { const long long now = clock64(); pc[k] += now - pt; pt = now; }
  const uint32_t barAddr = smemAddr(bar);
  const uint32_t stripBytes = uint32_t(g.stripStride) * 4u;
  if (tid == 0) {
    flags[0] = 0;
    mbarInit(barAddr, 1);
    fenceBarrierInit();
    mbarExpectTx(barAddr, stripBytes);
    const char* src = reinterpret_cast<const char*>(g.strips + size_t(b) * g.stripStride);
    for (uint32_t off = 0; off < stripBytes; off += 16384u) bulkLoad(smemAddr(strips) + off, src + off, stripBytes - off < 16384u ? stripBytes - off : 16384u, barAddr);
  }
  if (warp == 1) tmemAlloc(smemAddr(tmemSlot), uint32_t(tmemColumns));
  if (tid >= 64 && tid < 128) strips[g.stripStride + (tid - 64)] = 0.f; // the all-zero strip that pads odd pair lists
  tcgenFenceBeforeSync();
  __syncthreads();
  tcgenFenceAfterSync();
  const uint32_t tmemBase = *reinterpret_cast<volatile uint32_t*>(tmemSlot);
  const uint32_t tmemWarp = tmemBase + (uint32_t((warp & 3) * 32) << 16) + uint32_t((warp >> 2) * (tmemColumns / 2)); // lane quarter; two warps share one
  mbarWaitRelaxed(barAddr, 0);
  MB2_GC(0)
  const int32_t* __restrict__ gtab = g.blob;
  const int32_t* tileOrder = gtab + g.offTileOrder, *tileQuadStart = gtab + g.offTilePairStart, *quads = gtab + g.offPairA;
  const int32_t* colStripStart = gtab + g.offColStripStart, *colStrip = gtab + g.offColStrip, *stripRow = gtab + g.offStripRow, *tileInfo = gtab + g.offTileInfo;
  // ---- Gram ----
  {
    int slot = 0;
    for (int ti = warp; ti < g.numOrder; ti += kGramThreads / 32, ++slot) {
      const int t = tileOrder[ti];
      if (t < 0) continue;
      float acc[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      gramTileAccumulate(strips, quads, tileQuadStart[t], tileQuadStart[t + 1], lane, acc);
      tmemPark8(tmemWarp + 8u * slot, &acc[0][0]);
    }
    for (int K = hw; K < S.numTileCols; K += kGramThreads / 16)
      y[16 * K + hl] = gramVectorEntry(strips, resid, colStrip, stripRow, colStripStart[K], colStripStart[K + 1], hl);
    tmemParkWait();
  }
  __syncthreads();
  MB2_GC(1)
  {
    int slot = 0;
    for (int ti = warp; ti < g.numOrder; ti += kGramThreads / 32, ++slot) {
      const int t = tileOrder[ti];
      if (t < 0) continue;
      float acc[2][4];
      tmemFetch8(tmemWarp + 8u * slot, &acc[0][0]);
      gramTileStore(tiles + size_t(t) * 256, acc, tileInfo[t], g.regularization, lane);
    }
    for (int s2 = tid; s2 < S.nPad; s2 += kGramThreads) {
      const int p = S.perm[s2];
      if (p >= 0) gsub[p] = y[s2];
    }
  }
  __syncthreads();
  MB2_GC(2)
  // ---- level-scheduled Cholesky (same phases as choleskyScheduledKernel) ----
  for (int L = 0; L < S.numLevels; ++L) {
    for (int ci = S.levelColStart[L] + warp; ci < S.levelColStart[L + 1]; ci += kGramThreads / 32) {
      const int K = S.levelCols[ci];
      cholDiagTile(tiles + size_t(S.diagTile[K]) * 256, y + 16 * K, lane, c.regularization, flags);
    }
    __syncthreads();
    MB2_GC(3)
    for (int pi = S.levelPanelStart[L] + warp; pi < S.levelPanelStart[L + 1]; pi += kGramThreads / 32) {
      float* ptile = tiles + size_t(S.panelTile[pi]) * 256;
      float x[2][4];
      cholPanelProduct(ptile, tiles + size_t(S.panelDiag[pi]) * 256, lane, x);
      cholPanelStore(ptile, lane, x);
    }
    __syncthreads();
    MB2_GC(4)
    for (int oi = S.levelOrderStart8[L] + warp; oi < S.levelOrderStart8[L + 1]; oi += kGramThreads / 32) { const int ti = S.taskOrder8[oi]; if (ti >= 0) cholUpdateTask(tiles, S, ti, lane); }
    for (int vi = S.levelVTaskStart[L] + hw; vi < S.levelVTaskStart[L + 1]; vi += kGramThreads / 16) cholVectorTask(tiles, y, S, vi, hl);
    __syncthreads();
    MB2_GC(5)
  }
  for (int L = S.numLevels - 1; L >= 0; --L) {
    for (int ci = S.levelColStart[L] + warp; ci < S.levelColStart[L + 1]; ci += kGramThreads / 32) cholBackwardColumn(tiles, y, S, S.levelCols[ci], lane);
    __syncthreads();
  }
  MB2_GC(6)
  cholFinish(c, b, n, y, gsub, flags[0] != 0, S.pos); // y holds the step per elimination slot
  MB2_GC(7)
  if constexpr (kProfile) {
    // a block from the middle of the grid: its SM is in steady state (the other resident CTAs are at unrelated phases), unlike block 0,
    // whose whole first wave starts in lock step
    if (b == int(gridDim.x >> 1) && tid == 0 && a.phaseCycles != nullptr)
      for (int k = 0; k < 8; ++k) a.phaseCycles[k] += (unsigned long long)pc[k];
  }
#undef MB2_GC
  tcgenFenceBeforeSync();
  __syncthreads();
  if (warp == 1) tmemFree(tmemBase, uint32_t(tmemColumns));
}

cudaError_t launchGramCholesky(const GramCholArgs& a, const CholSchedDev& sched, bool profile, cudaStream_t stream) {
  const size_t smem = gramCholeskySmemBytes(a.g.stripStride, a.g.blobInts, a.c.ns, sched.nPad, sched.numTiles, sched.blobInts);
  if (smem > size_t(g_maxSmemOptin) || (a.g.stripStride & 3) != 0) return cudaErrorInvalidConfiguration;
  // TMEM: two warps share a lane quarter, each parks up to `rounds` tiles of 8 columns; allocations are powers of two >= 32,
  // and the (up to four) CTAs of an SM must fit in its 512 columns together
  const int rounds = std::max(a.g.numOrder / (kGramThreads / 32), 1);
  int columns = 32;
  while (columns < 2 * 8 * rounds) columns <<= 1;
  if (columns > 128) return cudaErrorInvalidConfiguration;
  cudaError_t e = profile ? cudaFuncSetAttribute(gramCholeskyKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))
                          : cudaFuncSetAttribute(gramCholeskyKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  e = profile ? cudaFuncSetAttribute(gramCholeskyKernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, int(cudaSharedmemCarveoutMaxShared))
              : cudaFuncSetAttribute(gramCholeskyKernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, int(cudaSharedmemCarveoutMaxShared));
  if (e != cudaSuccess) return e;
  if (profile) gramCholeskyKernel<true><<<a.c.batch, kGramThreads, smem, stream>>>(a, sched, columns);
  else gramCholeskyKernel<false><<<a.c.batch, kGramThreads, smem, stream>>>(a, sched, columns);
  return cudaGetLastError();
}

} // namespace mb2
#include "ik_qr.cuh"
#include "ik_tr_qr.cuh"
namespace mb2 {

cudaError_t initKernelAttributes() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&g_numSms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&g_maxSmemPerSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  if (e != cudaSuccess) return e;
  return cudaDeviceGetAttribute(&g_maxSmemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
}

// ------------------------------------------------------------------------------------------------
// Small element-wise kernels: line search (gauss_newton_solver.cpp:283-313,
// subset_gauss_newton_solver.cpp:119-141), SolverT bookkeeping, target normalisation.
// ------------------------------------------------------------------------------------------------
__global__ void trialUpdateKernel(int batch, const float* thetaOrig, int ldTheta, const float* delta, int ns, const int32_t* cols, const float* scale,
                                  float* thetaTrial, const int32_t* active) {
  const int b = blockIdx.x;
  if (active[b] == 0) return;
  const float s = scale[b];
  for (int i = threadIdx.x; i < ns; i += blockDim.x) {
    const int c = cols[i];
    if (c >= 0) thetaTrial[size_t(b) * ldTheta + c] = thetaOrig[size_t(b) * ldTheta + c] - s * delta[size_t(b) * ns + i]; // c < 0: alignment column
  }
}
cudaError_t launchTrialUpdate(int batch, const float* thetaOrig, int ldTheta, const float* delta, int ns, const int32_t* cols, const float* scale,
                              float* thetaTrial, const int32_t* active, cudaStream_t stream) {
  trialUpdateKernel<<<batch, 128, 0, stream>>>(batch, thetaOrig, ldTheta, delta, ns, cols, scale, thetaTrial, active);
  return cudaGetLastError();
}

__global__ void lineSearchStepKernel(const LineSearchArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.batch || a.searching[b] == 0) return;
  const double error = a.errors[b], errorNew = a.trialErrors[b];
  const float scale = a.scale[b];
  bool accept;
  if (!a.subsetVariant) {
    const float scaledError = 1e-3f * (float)error; // kC1 * error_ in T
    accept = (error - errorNew) >= (double)(scale * scaledError);
  } else {
    accept = (error - errorNew) >= (double)(1e-4f * scale) * (double)a.gradDotDelta[b];
  }
  if (accept || a.step >= 9) a.searching[b] = 0;
  else a.scale[b] = scale * 0.5f;
}
cudaError_t launchLineSearchStep(const LineSearchArgs& a, cudaStream_t stream) {
  lineSearchStepKernel<<<(a.batch + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

__global__ void commitTrialKernel(int batch, int numParams, int ldTheta, const float* thetaTrial, float* theta, const int32_t* active) {
  const int b = blockIdx.x;
  if (active[b] == 0) return;
  for (int i = threadIdx.x; i < numParams; i += blockDim.x) theta[size_t(b) * ldTheta + i] = thetaTrial[size_t(b) * ldTheta + i];
}
cudaError_t launchCommitTrial(int batch, int numParams, int ldTheta, const float* thetaTrial, float* theta, const int32_t* active, cudaStream_t stream) {
  commitTrialKernel<<<batch, 128, 0, stream>>>(batch, numParams, ldTheta, thetaTrial, theta, active);
  return cudaGetLastError();
}

__global__ void bookkeepingKernel(const BookkeepingArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.batch || a.active[b] == 0) return;
  const double error = a.errors[b], last = a.lastErrors[b];
  if (a.history != nullptr) a.history[size_t(b) * a.maxIterations + a.iteration] = error;
  const bool converged = fabs(last - error) / (fabs(error) + (double)FLT_MIN) <= (double)(a.threshold * FLT_EPSILON);
  a.iterations[b] = a.iteration + 1;
  if ((a.iteration >= a.minIterations && converged) || a.iteration + 1 >= a.maxIterations) a.active[b] = 0;
  else { a.lastErrors[b] = error; atomicAdd(a.activeCount, 1); }
}
cudaError_t launchBookkeeping(const BookkeepingArgs& a, cudaStream_t stream) {
  bookkeepingKernel<<<(a.batch + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

__global__ void normalizeQuatsKernel(float* base, int count, int strideFloats, int quatsPerRecord, int batch) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * quatsPerRecord) return;
  const int b = idx / quatsPerRecord, q = idx % quatsPerRecord;
  float* p = base + size_t(b) * strideFloats + 4 * q;
  const float n = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
  p[0] /= n; p[1] /= n; p[2] /= n; p[3] /= n;
}
// packed [batch][size] -> records [batch][strideFloats] at dst (a strided 2-D copy from the host takes one DMA descriptor per row)
__global__ void scatterTargetsKernel(const float* packed, float* dst, int size, int strideFloats, int batch) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= size_t(batch) * size) return;
  const size_t b = idx / size, k = idx % size;
  dst[b * strideFloats + k] = packed[idx];
}
cudaError_t launchScatterTargets(const float* packed, float* dst, int size, int strideFloats, int batch, cudaStream_t stream) {
  const size_t total = size_t(batch) * size;
  if (total == 0) return cudaSuccess;
  scatterTargetsKernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(packed, dst, size, strideFloats, batch);
  return cudaGetLastError();
}
cudaError_t launchNormalizeQuats(float* base, int count, int strideFloats, int quatsPerRecord, int batch, cudaStream_t stream) {
  const int total = batch * quatsPerRecord;
  if (total == 0) return cudaSuccess;
  normalizeQuatsKernel<<<(total + 127) / 128, 128, 0, stream>>>(base, count, strideFloats, quatsPerRecord, batch);
  return cudaGetLastError();
}

} // namespace mb2

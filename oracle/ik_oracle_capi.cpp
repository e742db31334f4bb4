// ORACLE — TEST INFRASTRUCTURE ONLY (see ik_oracle.hpp header). Flat C entry points for ctypes.
// All floating-point arrays cross this boundary as float64 and are narrowed to the solve type T
// (dtype 0 = float, 1 = double) inside, so one signature serves both precisions.
#include "ik_oracle.hpp"

#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>

#include <sched.h>
#include <cstdio>

using namespace oracle;

namespace {

struct FnHandle {
  int dtype;
  const Rig* rig;
  std::unique_ptr<SkeletonSolverFunction<float>> f;
  std::unique_ptr<SkeletonSolverFunction<double>> d;
  std::vector<uint8_t> enabled;
  bool enabledSet{false};
  double trustRegionRadius{1.0}; // TrustRegionQROptions::trustRegionRadius_
};

template <class T>
SkeletonSolverFunction<T>& get(FnHandle* h);
template <>
SkeletonSolverFunction<float>& get<float>(FnHandle* h) { return *h->f; }
template <>
SkeletonSolverFunction<double>& get<double>(FnHandle* h) { return *h->d; }

template <class T>
std::vector<T> narrow(const double* p, size_t n) {
  std::vector<T> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = T(p[i]);
  return v;
}

template <class T>
void normalizePlanes(ErrorFunction<T>& ef) { // PlaneDataT ctor: normal(inNormal.normalized()) (plane_error_function.h:35)
  for (size_t i = 0; i < ef.cparent.size(); ++i) {
    T* p = &ef.ctarget[4 * i];
    const T n = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    p[0] /= n; p[1] /= n; p[2] /= n;
  }
}

template <class T>
int addJointEf(FnHandle* h, int kind, double weight, double alpha, double c, int nc, const int* parents, const double* cw, const double* offsets, const double* targets) {
  ErrorFunction<T> ef;
  ef.kind = kind;
  ef.weight = T(weight);
  ef.lossAlpha = std::isinf(alpha) && alpha < 0 ? GeneralizedLoss<T>::kWelsch() : T(alpha);
  ef.lossC = T(c);
  const int per = kind == kPosition ? 3 : 4;                      // target record: 3 (position), 4 (orientation quaternion / plane normal + d)
  const int perOff = (kind == kPosition || kind == kPlane) ? 3 : 4; // offset record
  ef.cparent.assign(parents, parents + nc);
  ef.cweight.resize(nc);
  for (int i = 0; i < nc; ++i) ef.cweight[i] = float(cw[i]);
  ef.coffset = narrow<T>(offsets, size_t(nc) * perOff);
  ef.ctarget = narrow<T>(targets, size_t(nc) * per);
  if (kind == kPlane) normalizePlanes(ef);
  else if (kind != kPosition) { // OrientationDataT ctor normalises (orientation_error_function.h:33-35)
    for (int i = 0; i < nc; ++i) {
      for (std::vector<T>* arr : {&ef.coffset, &ef.ctarget}) {
        Quat<T> q{(*arr)[4 * i], (*arr)[4 * i + 1], (*arr)[4 * i + 2], (*arr)[4 * i + 3]};
        q = normalized(q);
        (*arr)[4 * i] = q.x; (*arr)[4 * i + 1] = q.y; (*arr)[4 * i + 2] = q.z; (*arr)[4 * i + 3] = q.w;
      }
    }
  }
  auto& fn = get<T>(h);
  fn.addErrorFunction(ef);
  if (h->enabledSet) fn.setEnabledParameters(h->enabled);
  return int(fn.errorFunctions.size()) - 1;
}

template <class T>
void setTargetsFn(SkeletonSolverFunction<T>& fn, int idx, const double* t) {
  auto& ef = fn.errorFunctions.at(idx);
  if (ef.kind == kState) {
    ef.targetState = narrow<T>(t, size_t(fn.rig->numJoints) * 8);
  } else if (ef.kind == kPosition) {
    ef.ctarget = narrow<T>(t, ef.cparent.size() * 3);
  } else if (ef.kind == kPlane) {
    ef.ctarget = narrow<T>(t, ef.cparent.size() * 4);
    normalizePlanes(ef);
  } else if (ef.kind == kModelParameters) {
    ef.targetParameters = narrow<T>(t, size_t(fn.rig->numParams));
  } else {
    ef.ctarget = narrow<T>(t, ef.cparent.size() * 4);
    for (size_t i = 0; i < ef.cparent.size(); ++i) {
      Quat<T> q{ef.ctarget[4 * i], ef.ctarget[4 * i + 1], ef.ctarget[4 * i + 2], ef.ctarget[4 * i + 3]};
      q = normalized(q);
      ef.ctarget[4 * i] = q.x; ef.ctarget[4 * i + 1] = q.y; ef.ctarget[4 * i + 2] = q.z; ef.ctarget[4 * i + 3] = q.w;
    }
  }
}

template <class T>
void setTargets(FnHandle* h, int idx, const double* t) { setTargetsFn<T>(get<T>(h), idx, t); }

template <class T>
int targetSize(FnHandle* h, int idx) {
  auto& ef = get<T>(h).errorFunctions.at(idx);
  if (ef.kind == kState) return h->rig->numJoints * 8;
  if (ef.kind == kLimit) return 0;
  if (ef.kind == kModelParameters) return h->rig->numParams;
  return int(ef.cparent.size()) * (ef.kind == kPosition ? 3 : 4);
}

struct SolveOpts {
  int64_t minIterations, maxIterations;
  double threshold, regularization;
  int doLineSearch, useBlockJtJ, subsetSolver;
  double trustRegionRadius{1.0};
};

template <class T>
double solveOne(SkeletonSolverFunction<T>& fn, const std::vector<uint8_t>* enabled, const SolveOpts& o, double* params, int* iters, double* hist) {
  GaussNewtonOptions go;
  go.minIterations = size_t(o.minIterations);
  go.maxIterations = size_t(o.maxIterations);
  go.threshold = float(o.threshold);
  go.regularization = float(o.regularization);
  go.doLineSearch = o.doLineSearch != 0;
  go.useBlockJtJ = o.useBlockJtJ != 0;
  go.subsetSolver = o.subsetSolver == 1; // 0 GaussNewtonSolverT, 1 SubsetGaussNewtonSolverT, 2 GaussNewtonSolverQRT, 3 TrustRegionQRT
  go.qrSolver = o.subsetSolver == 2;
  go.trustRegionQr = o.subsetSolver == 3;
  go.trustRegionRadius = float(o.trustRegionRadius);
  GaussNewtonSolver<T> solver(go, &fn);
  if (enabled) solver.setEnabledParameters(*enabled);
  std::vector<T> p = narrow<T>(params, fn.numParameters);
  const double err = solver.solve(p);
  for (int i = 0; i < fn.numParameters; ++i) params[i] = double(p[i]);
  if (iters) *iters = int(solver.errorHistory.size());
  if (hist) for (size_t i = 0; i < solver.errorHistory.size(); ++i) hist[i] = solver.errorHistory[i];
  return err;
}

template <class T>
void solveBatch(FnHandle* h, const SolveOpts& o, int B, double* params, const double* const* targets, int nthreads, double* errors, int* iters, double* finalErrors) {
  const auto& proto = get<T>(h);
  const int n = proto.numParameters;
  const int nef = int(proto.errorFunctions.size());
  std::vector<int> tsz(nef);
  for (int e = 0; e < nef; ++e) tsz[e] = targetSize<T>(h, e);
  std::atomic<int> next{0};
  auto work = [&]() {
    // one solver function + solver per thread, as pymomentum/tensor_ik/tensor_ik.cpp:127-162
    SkeletonSolverFunction<T> fn(proto);
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= B) break;
      for (int e = 0; e < nef; ++e)
        if (targets && targets[e] && tsz[e] > 0) setTargetsFn<T>(fn, e, targets[e] + size_t(b) * tsz[e]);
      errors[b] = solveOne<T>(fn, h->enabledSet ? &h->enabled : nullptr, o, params + size_t(b) * n, iters ? iters + b : nullptr, nullptr);
      if (finalErrors) {
        std::vector<T> p = narrow<T>(params + size_t(b) * n, n);
        finalErrors[b] = fn.getError(p.data());
      }
    }
  };
  if (nthreads <= 1) { work(); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work);
  for (auto& t : th) t.join();
}

} // namespace

#define DISPATCH(h, expr_f, expr_d) ((h)->dtype == 0 ? (expr_f) : (expr_d))

template <class T>
static int addStateEf(FnHandle* h, double weight, int rotErrType, double posWgt, double rotWgt, const double* posW, const double* rotW, const double* target) {
  ErrorFunction<T> ef;
  ef.kind = kState;
  ef.weight = T(weight);
  ef.rotErrType = rotErrType;
  ef.posWgt = T(float(posWgt)); // setWeights takes float (state_error_function.h:83-86)
  ef.rotWgt = T(float(rotWgt));
  ef.targetPosW = narrow<T>(posW, h->rig->numJoints);
  ef.targetRotW = narrow<T>(rotW, h->rig->numJoints);
  if (target) ef.targetState = narrow<T>(target, size_t(h->rig->numJoints) * 8);
  auto& f = get<T>(h);
  f.addErrorFunction(ef);
  if (h->enabledSet) f.setEnabledParameters(h->enabled);
  return int(f.errorFunctions.size()) - 1;
}
template <class T>
static int addLimitEf(FnHandle* h, double weight, double alpha, double c) {
  ErrorFunction<T> ef;
  ef.kind = kLimit;
  ef.weight = T(weight);
  ef.lossAlpha = std::isinf(alpha) && alpha < 0 ? GeneralizedLoss<T>::kWelsch() : T(alpha);
  ef.lossC = T(c);
  auto& f = get<T>(h);
  f.addErrorFunction(ef);
  if (h->enabledSet) f.setEnabledParameters(h->enabled);
  return int(f.errorFunctions.size()) - 1;
}
template <class T>
static double getErrorT(FnHandle* h, const double* params) {
  auto& f = get<T>(h);
  std::vector<T> p = narrow<T>(params, f.numParameters);
  return f.getError(p.data());
}
template <class T>
static int jacRowsT(FnHandle* h) {
  auto& f = get<T>(h);
  int total = 0;
  for (size_t i = 0; i < f.errorFunctions.size(); ++i) total += f.getJacobianBlockSize(i);
  return padToSimdAlignment(total);
}
template <class T>
static double getJacobianT(FnHandle* h, const double* params, double* jac, double* res, int* actualRows) {
  auto& f = get<T>(h);
  std::vector<T> p = narrow<T>(params, f.numParameters);
  Mat<T> J;
  std::vector<T> r;
  int rows = 0;
  const double e = f.getJacobian(p.data(), J, r, rows);
  for (size_t i = 0; i < J.a.size(); ++i) jac[i] = double(J.a[i]);
  for (size_t i = 0; i < r.size(); ++i) res[i] = double(r[i]);
  *actualRows = rows;
  return e;
}
template <class T>
static double getJtJRT(FnHandle* h, const double* params, double* jtj, double* jtr) {
  auto& f = get<T>(h);
  std::vector<T> p = narrow<T>(params, f.numParameters);
  Mat<T> H;
  std::vector<T> g;
  const double e = f.getJtJR(p.data(), H, g);
  for (size_t i = 0; i < H.a.size(); ++i) jtj[i] = double(H.a[i]);
  for (size_t i = 0; i < g.size(); ++i) jtr[i] = double(g[i]);
  return e;
}
template <class T>
static void fkT(FnHandle* h, const double* params, double* xf, double* rotAxis, double* transAxis) {
  auto& f = get<T>(h);
  std::vector<T> p = narrow<T>(params, f.numParameters);
  f.updateState(p.data());
  for (int j = 0; j < h->rig->numJoints; ++j) {
    const auto& js = f.state.jointState[j];
    double* o = xf + 8 * j;
    o[0] = js.world.t.x; o[1] = js.world.t.y; o[2] = js.world.t.z;
    o[3] = js.world.q.x; o[4] = js.world.q.y; o[5] = js.world.q.z; o[6] = js.world.q.w;
    o[7] = js.world.s;
    if (rotAxis) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) rotAxis[9 * j + 3 * c + r] = js.rotationAxis(r, c);
    if (transAxis) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) transAxis[9 * j + 3 * c + r] = js.translationAxis(r, c);
  }
}

extern "C" {

void* orc_rig_create(int J, const int* parents, const double* offsets, const double* prerot, int n, const int* outer, const int* inner, const double* vals, const double* ptoffsets) {
  Rig* r = new Rig;
  r->numJoints = J;
  r->parent.assign(parents, parents + J);
  r->offset.resize(3 * J);
  r->prerot.resize(4 * J);
  for (int i = 0; i < 3 * J; ++i) r->offset[i] = float(offsets[i]);
  for (int i = 0; i < 4 * J; ++i) r->prerot[i] = float(prerot[i]);
  r->numParams = n;
  const int rows = J * kParametersPerJoint;
  r->outer.assign(outer, outer + rows + 1);
  const int nnz = outer[rows];
  r->inner.assign(inner, inner + nnz);
  r->vals.resize(nnz);
  for (int i = 0; i < nnz; ++i) r->vals[i] = float(vals[i]);
  r->ptOffsets.resize(rows);
  for (int i = 0; i < rows; ++i) r->ptOffsets[i] = float(ptoffsets[i]);
  std::vector<uint8_t> all(n, 1);
  r->activeJointParamsDefault = r->computeActiveJointParams(all);
  return r;
}
void orc_rig_add_limit(void* rig, int type, double weight, const int* i4, const double* f27) {
  Limit l;
  l.type = type;
  l.weight = float(weight);
  for (int k = 0; k < 4; ++k) l.i[k] = i4[k];
  for (int k = 0; k < 27; ++k) l.f[k] = float(f27[k]);
  static_cast<Rig*>(rig)->limits.push_back(l);
}
void orc_rig_destroy(void* rig) { delete static_cast<Rig*>(rig); }

void* orc_fn_create(void* rig, int dtype) {
  FnHandle* h = new FnHandle;
  h->dtype = dtype;
  h->rig = static_cast<Rig*>(rig);
  if (dtype == 0) h->f = std::make_unique<SkeletonSolverFunction<float>>(h->rig);
  else h->d = std::make_unique<SkeletonSolverFunction<double>>(h->rig);
  return h;
}
void orc_fn_destroy(void* fn) { delete static_cast<FnHandle*>(fn); }

int orc_fn_add_joint_ef(void* fn, int kind, double weight, double alpha, double c, int nc, const int* parents, const double* cw, const double* offsets, const double* targets) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  return DISPATCH(h, addJointEf<float>(h, kind, weight, alpha, c, nc, parents, cw, offsets, targets), addJointEf<double>(h, kind, weight, alpha, c, nc, parents, cw, offsets, targets));
}

int orc_fn_add_state_ef(void* fn, double weight, int rotErrType, double posWgt, double rotWgt, const double* posW, const double* rotW, const double* target) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  return DISPATCH(h, addStateEf<float>(h, weight, rotErrType, posWgt, rotWgt, posW, rotW, target), addStateEf<double>(h, weight, rotErrType, posWgt, rotWgt, posW, rotW, target));
}

int orc_fn_set_half_plane(void* fn, int idx, int above) { // PlaneErrorFunctionT(character, above)
  FnHandle* h = static_cast<FnHandle*>(fn);
  if (h->dtype == 0) h->f->errorFunctions.at(idx).halfPlane = above != 0; else h->d->errorFunctions.at(idx).halfPlane = above != 0;
  return 0;
}
int orc_fn_add_model_parameters_ef(void* fn, double weight, const double* targetParameters, const double* targetWeights) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  auto add = [&](auto& f, auto tag) {
    using T = decltype(tag);
    ErrorFunction<T> ef;
    ef.kind = kModelParameters;
    ef.weight = T(weight);
    ef.targetParameters = narrow<T>(targetParameters, size_t(h->rig->numParams));
    ef.targetWeights = narrow<T>(targetWeights, size_t(h->rig->numParams));
    f.addErrorFunction(ef);
    if (h->enabledSet) f.setEnabledParameters(h->enabled);
    return int(f.errorFunctions.size()) - 1;
  };
  return h->dtype == 0 ? add(*h->f, float()) : add(*h->d, double());
}

int orc_fn_add_limit_ef(void* fn, double weight, double alpha, double c) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  return DISPATCH(h, addLimitEf<float>(h, weight, alpha, c), addLimitEf<double>(h, weight, alpha, c));
}

void orc_fn_set_targets(void* fn, int idx, const double* t) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  if (h->dtype == 0) setTargets<float>(h, idx, t); else setTargets<double>(h, idx, t);
}
void orc_fn_set_cweights(void* fn, int idx, const double* w) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  if (h->dtype == 0) { auto& ef = h->f->errorFunctions.at(idx); for (size_t i = 0; i < ef.cweight.size(); ++i) ef.cweight[i] = float(w[i]); }
  else { auto& ef = h->d->errorFunctions.at(idx); for (size_t i = 0; i < ef.cweight.size(); ++i) ef.cweight[i] = float(w[i]); }
}
void orc_fn_set_weight(void* fn, int idx, double w) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  if (h->dtype == 0) h->f->errorFunctions.at(idx).weight = float(w); else h->d->errorFunctions.at(idx).weight = w;
}
void orc_fn_set_enabled(void* fn, const uint8_t* enabled) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  h->enabled.assign(enabled, enabled + h->rig->numParams);
  h->enabledSet = true;
  if (h->dtype == 0) h->f->setEnabledParameters(h->enabled); else h->d->setEnabledParameters(h->enabled);
}
int orc_fn_actual_parameters(void* fn) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  return DISPATCH(h, h->f->actualParameters, h->d->actualParameters);
}

double orc_fn_get_error(void* fn, const double* params) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  return DISPATCH(h, getErrorT<float>(h, params), getErrorT<double>(h, params));
}

int orc_fn_jacobian_rows(void* fn) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  return DISPATCH(h, jacRowsT<float>(h), jacRowsT<double>(h));
}

double orc_fn_get_jacobian(void* fn, const double* params, double* jac, double* res, int* actualRows) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  return DISPATCH(h, getJacobianT<float>(h, params, jac, res, actualRows), getJacobianT<double>(h, params, jac, res, actualRows));
}

double orc_fn_get_jtjr(void* fn, const double* params, double* jtj, double* jtr) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  return DISPATCH(h, getJtJRT<float>(h, params, jtj, jtr), getJtJRT<double>(h, params, jtj, jtr));
}

// Forward kinematics: xf[J][8] = (t, q xyzw, s); axes column-major 3x3 per joint.
void orc_fn_fk(void* fn, const double* params, double* xf, double* rotAxis, double* transAxis) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  if (h->dtype == 0) fkT<float>(h, params, xf, rotAxis, transAxis); else fkT<double>(h, params, xf, rotAxis, transAxis);
}

void orc_fn_set_trust_region_radius(void* fn, double radius) { static_cast<FnHandle*>(fn)->trustRegionRadius = radius; }

double orc_solve(void* fn, int64_t minIt, int64_t maxIt, double threshold, double regularization, int doLineSearch, int useBlockJtJ, int subsetSolver, double* params, int* iters, double* errHistory) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  SolveOpts o{minIt, maxIt, threshold, regularization, doLineSearch, useBlockJtJ, subsetSolver};
  o.trustRegionRadius = h->trustRegionRadius;
  const std::vector<uint8_t>* en = h->enabledSet ? &h->enabled : nullptr;
  return DISPATCH(h, solveOne<float>(*h->f, en, o, params, iters, errHistory), solveOne<double>(*h->d, en, o, params, iters, errHistory));
}

// Batched solve, one solver per instance, instances distributed over `nthreads` host threads
// (pymomentum/tensor_ik/tensor_ik.cpp:127). targets[e] = pointer to [B x targetSize(e)] or NULL (shared).
// Returns wall seconds of the threaded region.
double orc_solve_batch(void* fn, int64_t minIt, int64_t maxIt, double threshold, double regularization, int doLineSearch, int useBlockJtJ, int subsetSolver, int B, double* params, const double* const* targets, int nthreads, double* errors, int* iters, double* finalErrors) {
  FnHandle* h = static_cast<FnHandle*>(fn);
  SolveOpts o{minIt, maxIt, threshold, regularization, doLineSearch, useBlockJtJ, subsetSolver};
  const auto t0 = std::chrono::steady_clock::now();
  if (h->dtype == 0) solveBatch<float>(h, o, B, params, targets, nthreads, errors, iters, finalErrors);
  else solveBatch<double>(h, o, B, params, targets, nthreads, errors, iters, finalErrors);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// OnlineHouseholderQR on its own (known-answer tests of math/online_householder_qr): A row-major [rows x n], added in `numChunks`
// consecutive row chunks; x = result(), atb = At_times_b().
void orc_online_qr(int dtype, int rows, int n, double lambda, const double* A, const double* b, int numChunks, const int* chunkRows, double* x, double* atb) {
  auto run = [&](auto tag) {
    using T = decltype(tag);
    OnlineHouseholderQR<T> qr;
    qr.reset(n, T(lambda));
    int r0 = 0;
    for (int c = 0; c < numChunks; ++c) {
      const int p = chunkRows[c];
      std::vector<T> Ac(size_t(p) * n), bc(p);
      for (int k = 0; k < p; ++k) { bc[k] = T(b[r0 + k]); for (int j = 0; j < n; ++j) Ac[size_t(j) * p + k] = T(A[size_t(r0 + k) * n + j]); }
      qr.addMutating(Ac.data(), p, p, bc.data());
      r0 += p;
    }
    const std::vector<T> xs = qr.result(), g = qr.AtTimesB();
    for (int j = 0; j < n; ++j) { x[j] = double(xs[j]); atb[j] = double(g[j]); }
  };
  if (dtype == 0) run(float(0)); else run(double(0));
  (void)rows;
}

// Host threads this process may actually use: the CPU affinity mask, capped by the cgroup CPU quota (cpu.max = "<quota> <period>"),
// not std::thread::hardware_concurrency() (which reports every core of the host even inside a container with a few CPUs' worth of quota).
int orc_hardware_threads() {
  int n = int(std::thread::hardware_concurrency());
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0) { const int c = CPU_COUNT(&set); if (c > 0) n = c; }
  for (const char* path : {"/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"}) {
    FILE* f = std::fopen(path, "r");
    if (!f) continue;
    char buf[128] = {0};
    if (std::fgets(buf, sizeof(buf), f)) {
      long long quota = -1, period = 100000;
      if (std::strncmp(buf, "max", 3) != 0) {
        if (std::sscanf(buf, "%lld %lld", &quota, &period) < 2) {
          period = 100000;
          FILE* pf = std::fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r");
          if (pf) { if (std::fscanf(pf, "%lld", &period) != 1) period = 100000; std::fclose(pf); }
        }
        if (quota > 0 && period > 0) { const int q = int((quota + period - 1) / period); if (q > 0 && q < n) n = q; }
      }
    }
    std::fclose(f);
    break;
  }
  return n > 0 ? n : 1;
}

} // extern "C"

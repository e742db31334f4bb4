// Mixed-rig batches (BASELINE.json configs[4]): a host component above the batched C-ABI that takes a heterogeneous
// list of IK instances (rig, position constraints, initial parameters), sorts it into buckets that can share one solver plan,
// runs every bucket as one batched solve on its own stream and scatters the results back in input order.
//
// The reference handles such a batch one element at a time (pymomentum/tensor_ik/tensor_ik.cpp:127-177: per element, build the
// error functions, a solver function and a solver). On the device the sparsity pattern of the normal equations is planned once
// per (rig, constraint parents), so that is the bucket key; inside a bucket the offsets, targets and weights are per instance and
// an instance with fewer constraints is padded with zero-weight rows (joint_error_function-inl.h:197-199 skips them).
#include "../../include/momentum_b200.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace {

thread_local std::string g_mixedError;
int mixedFail(int code, const std::string& msg) {
  g_mixedError = msg;
  return code;
}

struct Instance {
  int rig{0};
  std::vector<int32_t> parents;
  std::vector<float> offsets, weights, targets, theta;
  int bucket{-1}, slot{-1};
  double error{0};
  int32_t iterations{0}, status{0};
};

struct Bucket {
  int rig{0};
  std::vector<int32_t> parents; // of its longest instance; every member's parents are a prefix
  std::vector<int> members;
  mb2_solver_function* fn{nullptr};
  mb2_solver* solver{nullptr};
  std::vector<float> theta;     // [members][n], bucket order
  uint64_t iterations{0};
  ~Bucket() {
    if (solver) mb2_solver_destroy(solver);
    if (fn) mb2_solver_function_destroy(fn);
  }
};

} // namespace

struct mb2_mixed_batch {
  int device{0};
  int granule{8};
  bool useLimits{false};
  float limitWeight{1.f};
  std::vector<const mb2_character*> rigs;
  std::vector<int> rigParams;
  std::vector<Instance> instances;
  std::vector<std::unique_ptr<Bucket>> buckets;
  bool planned{false};
  int64_t realRows{0}, paddedRows{0};
};

extern "C" {

const char* mb2_mixed_batch_last_error(void) { return g_mixedError.empty() ? mb2_last_error() : g_mixedError.c_str(); }

int mb2_mixed_batch_create(int device, int32_t granule, mb2_mixed_batch** out) {
  if (!out) return mixedFail(MB2_ERR_INVALID_ARGUMENT, "null argument");
  auto b = std::make_unique<mb2_mixed_batch>();
  b->device = device;
  b->granule = granule > 0 ? granule : 8;
  *out = b.release();
  return MB2_OK;
}
void mb2_mixed_batch_destroy(mb2_mixed_batch* b) { delete b; }

int mb2_mixed_batch_add_rig(mb2_mixed_batch* b, const mb2_character* c, int32_t num_parameters, int32_t* rig_id) {
  if (!b || !c || num_parameters <= 0) return mixedFail(MB2_ERR_INVALID_ARGUMENT, "invalid rig");
  b->rigs.push_back(c);
  b->rigParams.push_back(num_parameters);
  if (rig_id) *rig_id = int32_t(b->rigs.size()) - 1;
  return MB2_OK;
}

int mb2_mixed_batch_use_limits(mb2_mixed_batch* b, int32_t enabled, float weight) {
  if (!b) return mixedFail(MB2_ERR_INVALID_ARGUMENT, "null batch");
  b->useLimits = enabled != 0;
  b->limitWeight = weight;
  b->planned = false;
  return MB2_OK;
}

int mb2_mixed_batch_add_instance(mb2_mixed_batch* b, int32_t rig_id, int32_t nc, const int32_t* parents, const float* offsets, const float* weights, const float* targets,
                                 const float* theta0, int32_t* instance_id) {
  if (!b || rig_id < 0 || rig_id >= int32_t(b->rigs.size()) || nc < 0 || (nc > 0 && (!parents || !offsets || !weights || !targets)) || !theta0)
    return mixedFail(MB2_ERR_INVALID_ARGUMENT, "invalid instance");
  Instance in;
  in.rig = rig_id;
  in.parents.assign(parents, parents + nc);
  in.offsets.assign(offsets, offsets + 3 * size_t(nc));
  in.weights.assign(weights, weights + nc);
  in.targets.assign(targets, targets + 3 * size_t(nc));
  in.theta.assign(theta0, theta0 + b->rigParams[rig_id]);
  b->instances.push_back(std::move(in));
  b->planned = false;
  if (instance_id) *instance_id = int32_t(b->instances.size()) - 1;
  return MB2_OK;
}

// Buckets: per rig, instances in decreasing constraint count; an instance joins the first bucket whose parent list starts with its
// own and is at most `granule - 1` constraints longer (bounded padding), else it opens a bucket of its own.
static int planBuckets(mb2_mixed_batch* b) {
  b->buckets.clear();
  b->realRows = b->paddedRows = 0;
  std::vector<int> order(b->instances.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = int(i);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
    const Instance &a = b->instances[x], &c = b->instances[y];
    if (a.rig != c.rig) return a.rig < c.rig;
    if (a.parents.size() != c.parents.size()) return a.parents.size() > c.parents.size();
    return a.parents < c.parents;
  });
  // (rig, rounded-up count) -> candidate buckets
  std::map<std::pair<int, int>, std::vector<int>> index;
  for (int id : order) {
    Instance& in = b->instances[id];
    const int c = int(in.parents.size());
    const int cls = (c + b->granule - 1) / b->granule;
    int found = -1;
    auto it = index.find({in.rig, cls});
    if (it != index.end())
      for (int bi : it->second) {
        const Bucket& bk = *b->buckets[bi];
        if (int(bk.parents.size()) >= c && std::equal(in.parents.begin(), in.parents.end(), bk.parents.begin())) { found = bi; break; }
      }
    if (found < 0) {
      auto bk = std::make_unique<Bucket>();
      bk->rig = in.rig;
      bk->parents = in.parents;
      b->buckets.push_back(std::move(bk));
      found = int(b->buckets.size()) - 1;
      index[{in.rig, cls}].push_back(found);
    }
    in.bucket = found;
    in.slot = int(b->buckets[found]->members.size());
    b->buckets[found]->members.push_back(id);
    b->realRows += 3 * c;
    b->paddedRows += 3 * int64_t(b->buckets[found]->parents.size());
  }
  b->planned = true;
  return MB2_OK;
}

int mb2_mixed_batch_solve(mb2_mixed_batch* b, const mb2_gauss_newton_options* opt) {
  if (!b || !opt) return mixedFail(MB2_ERR_INVALID_ARGUMENT, "null argument");
  g_mixedError.clear();
  if (!b->planned) {
    const int rc = planBuckets(b);
    if (rc != MB2_OK) return rc;
  }
  // build (once) and launch every bucket; the launches are asynchronous, buckets overlap on the device
  for (auto& up : b->buckets) {
    Bucket& bk = *up;
    const int B = int(bk.members.size()), nc = int(bk.parents.size()), n = b->rigParams[bk.rig];
    int rc;
    if (!bk.fn) {
      if ((rc = mb2_solver_function_create(b->rigs[bk.rig], B, &bk.fn)) != MB2_OK) return rc;
      std::vector<float> ones(size_t(std::max(nc, 1)), 1.f);
      int32_t idx = -1;
      if ((rc = mb2_add_position_error_function_instanced(bk.fn, 1.f, 2.f, 1.f, nc, bk.parents.data(), ones.data(), &idx)) != MB2_OK) return rc;
      if (b->useLimits && (rc = mb2_add_limit_error_function(bk.fn, b->limitWeight, 2.f, 1.f, nullptr)) != MB2_OK) return rc;
      if ((rc = mb2_solver_create(bk.fn, opt, &bk.solver)) != MB2_OK) return rc;
    } else if ((rc = mb2_solver_set_options(bk.solver, opt)) != MB2_OK) {
      return rc;
    }
    // per-instance records: targets + offsets, weights (zero beyond the instance's own constraints), initial parameters
    std::vector<float> rec(size_t(B) * 6 * nc, 0.f), w(size_t(B) * nc, 0.f);
    bk.theta.assign(size_t(B) * n, 0.f);
    for (int s = 0; s < B; ++s) {
      const Instance& in = b->instances[bk.members[s]];
      const int c = int(in.parents.size());
      for (int k = 0; k < c; ++k) {
        for (int d = 0; d < 3; ++d) { rec[(size_t(s) * nc + k) * 6 + d] = in.targets[3 * k + d]; rec[(size_t(s) * nc + k) * 6 + 3 + d] = in.offsets[3 * k + d]; }
        w[size_t(s) * nc + k] = in.weights[k];
      }
      std::copy(in.theta.begin(), in.theta.end(), bk.theta.begin() + size_t(s) * n);
    }
    if (nc > 0) {
      if ((rc = mb2_set_constraint_weights(bk.fn, 0, w.data(), 1)) != MB2_OK) return rc;
      if ((rc = mb2_set_targets(bk.fn, 0, rec.data())) != MB2_OK) return rc;
    }
  }
  // every bucket is queued on its own handle's stream (H2D of the parameters, solve, D2H), then collected: buckets overlap on the device
  for (auto& up : b->buckets) {
    const int rc = mb2_solver_solve_async(up->solver, up->theta.data());
    if (rc != MB2_OK) return rc;
  }
  for (auto& up : b->buckets) {
    Bucket& bk = *up;
    const int B = int(bk.members.size());
    std::vector<double> err(B);
    std::vector<int32_t> its(B), st(B);
    const int rc = mb2_solver_wait(bk.solver, err.data(), its.data(), st.data());
    if (rc != MB2_OK) return rc;
    const int n = b->rigParams[bk.rig];
    bk.iterations = 0;
    for (int s = 0; s < B; ++s) {
      Instance& in = b->instances[bk.members[s]];
      std::copy(bk.theta.begin() + size_t(s) * n, bk.theta.begin() + size_t(s + 1) * n, in.theta.begin());
      in.error = err[s];
      in.iterations = its[s];
      in.status = st[s];
      bk.iterations += uint64_t(its[s]);
    }
  }
  return MB2_OK;
}

int mb2_mixed_batch_get_result(mb2_mixed_batch* b, int32_t instance_id, float* theta, double* error, int32_t* iterations, int32_t* status) {
  if (!b || instance_id < 0 || instance_id >= int32_t(b->instances.size())) return mixedFail(MB2_ERR_INVALID_ARGUMENT, "instance id out of range");
  const Instance& in = b->instances[instance_id];
  if (theta) std::copy(in.theta.begin(), in.theta.end(), theta);
  if (error) *error = in.error;
  if (iterations) *iterations = in.iterations;
  if (status) *status = in.status;
  return MB2_OK;
}

// every instance at once, in input order: theta concatenated (instance i at theta_offsets[i], its rig's parameter count long)
int mb2_mixed_batch_get_results(mb2_mixed_batch* b, float* theta, const int64_t* theta_offsets, double* errors, int32_t* iterations, int32_t* status) {
  if (!b) return mixedFail(MB2_ERR_INVALID_ARGUMENT, "null batch");
  for (size_t i = 0; i < b->instances.size(); ++i) {
    const Instance& in = b->instances[i];
    if (theta && theta_offsets) std::copy(in.theta.begin(), in.theta.end(), theta + theta_offsets[i]);
    if (errors) errors[i] = in.error;
    if (iterations) iterations[i] = in.iterations;
    if (status) status[i] = in.status;
  }
  return MB2_OK;
}

// re-arm for another solve of the same instances from new initial parameters (bucket handles and plans are kept)
int mb2_mixed_batch_set_parameters(mb2_mixed_batch* b, int32_t instance_id, const float* theta0) {
  if (!b || instance_id < 0 || instance_id >= int32_t(b->instances.size()) || !theta0) return mixedFail(MB2_ERR_INVALID_ARGUMENT, "invalid argument");
  Instance& in = b->instances[instance_id];
  std::copy(theta0, theta0 + in.theta.size(), in.theta.begin());
  return MB2_OK;
}

// stats: [0] instances, [1] buckets, [2] residual rows of the instances' own constraints, [3] rows after padding to the bucket's length,
// [4] largest bucket, [5] singleton buckets
int mb2_mixed_batch_stats(mb2_mixed_batch* b, int64_t stats[6]) {
  if (!b || !stats) return mixedFail(MB2_ERR_INVALID_ARGUMENT, "null argument");
  if (!b->planned) planBuckets(b);
  stats[0] = int64_t(b->instances.size());
  stats[1] = int64_t(b->buckets.size());
  stats[2] = b->realRows;
  stats[3] = b->paddedRows;
  int64_t largest = 0, single = 0;
  for (const auto& bk : b->buckets) { largest = std::max<int64_t>(largest, int64_t(bk->members.size())); single += bk->members.size() == 1 ? 1 : 0; }
  stats[4] = largest;
  stats[5] = single;
  return MB2_OK;
}

// per bucket: [0] rig id, [1] instances, [2] constraints (padded length), [3] Gauss-Newton iterations executed by its last solve
int mb2_mixed_batch_bucket_info(mb2_mixed_batch* b, int32_t bucket, int64_t info[4]) {
  if (!b || !info) return mixedFail(MB2_ERR_INVALID_ARGUMENT, "null argument");
  if (!b->planned) planBuckets(b);
  if (bucket < 0 || bucket >= int32_t(b->buckets.size())) return mixedFail(MB2_ERR_INVALID_ARGUMENT, "bucket index out of range");
  const Bucket& bk = *b->buckets[bucket];
  info[0] = bk.rig;
  info[1] = int64_t(bk.members.size());
  info[2] = int64_t(bk.parents.size());
  info[3] = int64_t(bk.iterations);
  return MB2_OK;
}

} // extern "C"

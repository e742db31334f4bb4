/*
 * momentum_b200 — C-ABI of the B200-native batched Gauss-Newton IK path.
 *
 * This is the drop-in boundary for ONE hot path of facebookresearch/momentum: the per-iteration
 * FK sweep -> residual/Jacobian (Position / Orientation / State / Limit) -> JtJ, Jtr -> damped
 * Cholesky -> parameter update, for a BATCH of independent IK instances that share one rig and one
 * constraint topology (the body of the dispenso::parallel_for at
 * pymomentum/tensor_ik/tensor_ik.cpp:127-177). Every entry point below names the reference
 * interface it stands in for (paths relative to the reference's momentum/ directory).
 *
 * Conventions
 *  - plain C types only; no C++/torch types cross this boundary.
 *  - every function returns MB2_OK (0) or an error code; mb2_last_error() gives the message of the
 *    last failure on the calling thread (reference: MT_CHECK/MT_THROW -> std::runtime_error,
 *    common/checks.h:36, common/exception.h:31; adapters rethrow).
 *  - host pointers unless the name says _device. Host inputs are copied during the call and never
 *    retained (reference ownership: the solver holds non-owning pointers, solver/solver.h:106).
 *  - handles are not thread-safe; one handle = one CUDA stream (reference: one solver + function per
 *    thread, tensor_ik.cpp:127-162; mutable scratch in skeleton_solver_function.h:89).
 *  - quaternions are (x, y, z, w) like Eigen::Quaternion::coeffs().
 *  - batched arrays are instance-major: [B][...].
 *  - there is no CPU fallback: every compute entry point fails with MB2_ERR_CUDA when no sm_100
 *    device is usable.
 */
#ifndef MOMENTUM_B200_H_
#define MOMENTUM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB2_PARAMETERS_PER_JOINT 7   /* character/types.h:21 kParametersPerJoint */
#define MB2_MAX_MODEL_PARAMETERS 2048 /* math/types.h:426-429 ParameterSet = std::bitset<2048> */
#define MB2_PARAMETER_SET_WORDS 32   /* 2048 / 64 */

typedef enum mb2_status {
  MB2_OK = 0,
  MB2_ERR_INVALID_ARGUMENT = 1, /* MT_CHECK failure in the reference */
  MB2_ERR_CUDA = 2,             /* CUDA runtime error / no usable device (no CPU fallback) */
  MB2_ERR_UNSUPPORTED = 3
} mb2_status;

/* Per-instance result status of a batched solve (reference: NaN/Inf guard of the batched caller,
 * tensor_ik.cpp:168-173; LLT info is ignored by the dense solver, gauss_newton_solver.cpp:251). */
typedef enum mb2_instance_status {
  MB2_INSTANCE_OK = 0,
  MB2_INSTANCE_CHOLESKY_BREAKDOWN = 1, /* a non-positive pivot was met (Eigen: NumericalIssue) */
  MB2_INSTANCE_NON_FINITE = 2          /* parameters became NaN/Inf */
} mb2_instance_status;

/* character/parameter_limits.h:20-33 LimitType */
typedef enum mb2_limit_type {
  MB2_LIMIT_MINMAX = 0,
  MB2_LIMIT_MINMAX_JOINT = 1,
  MB2_LIMIT_MINMAX_JOINT_PASSIVE = 2,
  MB2_LIMIT_LINEAR = 3,
  MB2_LIMIT_LINEAR_JOINT = 4,
  MB2_LIMIT_ELLIPSOID = 5,
  MB2_LIMIT_HALFPLANE = 6
} mb2_limit_type;

/* character/parameter_limits.h:117-127 ParameterLimit, flattened.
 *  MinMax:      i[0]=parameterIndex, f[0..1]=limits
 *  MinMaxJoint: i[0]=jointIndex, i[1]=jointParameter, f[0..1]=limits
 *  Linear:      i[0]=referenceIndex, i[1]=targetIndex, f[0]=scale, f[1]=offset, f[2]=rangeMin, f[3]=rangeMax
 *  LinearJoint: i[0..1]=reference joint/param, i[2..3]=target joint/param, f[0..3] as Linear
 *  HalfPlane:   i[0]=param1, i[1]=param2, f[0..1]=normal, f[2]=offset
 *  Ellipsoid:   i[0]=ellipsoidParent, i[1]=parent, f[0..11]=ellipsoid 3x4 row-major, f[12..23]=ellipsoidInv, f[24..26]=offset */
typedef struct mb2_parameter_limit {
  int32_t type;
  float weight;
  int32_t i[4];
  float f[27];
} mb2_parameter_limit;

/* state_error_function.h:17-32 RotationErrorType */
typedef enum mb2_rotation_error_type {
  MB2_ROTATION_MATRIX_DIFFERENCE = 0,
  MB2_QUATERNION_LOG_MAP = 1
} mb2_rotation_error_type;

/* How JtJ is formed on the device (extension; the reference always uses Eigen fp32/fp64 GEMM). */
typedef enum mb2_jtj_mode {
  MB2_JTJ_AUTO = 0,      /* tile-sparse Gram (mma.sync, three-term TF32 split: fp32-class, not bit-exact fp32) with the tile-scheduled Cholesky; else tcgen05 3xTF32 where the shape allows, else FP32 SIMT */
  MB2_JTJ_FP32_SIMT = 1, /* CUDA-core fp32 (validation path) */
  MB2_JTJ_TF32X3 = 2,    /* tcgen05 kind::tf32, 3-term split, fp32 accumulate in TMEM (fp32-class accuracy); up to 511 Jacobian columns */
  MB2_JTJ_TF32 = 3,      /* tcgen05 kind::tf32 single pass (~1e-3 relative; changes the GN path, not the fixed point) */
  MB2_JTJ_SPARSE_TILES = 4 /* tensor cores (mma.sync m16n8k8, hi*hi + hi*lo + lo*hi TF32 split, the lo*lo term is dropped: ~2^-21 relative) over the non-zero
                              strips of J only, straight into the Cholesky tile layout (tile-scheduled Cholesky only) */
} mb2_jtj_mode;

/* How (JtJ + lambda I) delta = Jtr is solved on the device (extension; the reference always runs a dense
 * Eigen::LLT, gauss_newton_solver.cpp:251). All modes solve the same system; they differ in elimination
 * order (rounding) and in what happens on a non-positive pivot: the dense Eigen-structured kernel
 * reproduces Eigen's early exit, the tile schedules substitute the damping for the pivot; both flag the
 * instance MB2_INSTANCE_CHOLESKY_BREAKDOWN. */
typedef enum mb2_cholesky_mode {
  MB2_CHOLESKY_AUTO = 0,            /* tile schedule on the sparsity pattern for >= 48 unknowns, else dense Eigen-structured */
  MB2_CHOLESKY_DENSE_EIGEN = 1,     /* blocked LLT with Eigen's block structure and failure semantics */
  MB2_CHOLESKY_TILES_DENSE = 2,     /* level-scheduled 16x16 tiles, every tile present */
  MB2_CHOLESKY_TILES_SPARSE = 3     /* level-scheduled tiles over the kinematic-tree sparsity of JtJ (min-degree order) */
} mb2_cholesky_mode;

/* Which kernels run one Gauss-Newton iteration on the tile path (extension; every mode solves the same system with the same device
 * functions, results agree to float rounding).
 *   GRAM_CHOLESKY  two launches per iteration: FK / residual / Jacobian strips, then ONE kernel that forms the stored tiles of
 *                  J^T J + lambda I from the strips on the tensor cores, parks them in TMEM, writes them over the dead strips in shared
 *                  memory and runs the tile Cholesky + update on them: the normal equations never exist in HBM.
 *   OFF            three launches per iteration (sweep, tile-sparse Gram to HBM, tile Cholesky).
 *   PERSISTENT     ONE launch per solve: groups of 256 threads keep an instance in shared memory for all of its iterations (FK sweep
 *                  included) and stop it on the device; no host round trip, no HBM traffic beyond theta / targets / results. Needs no
 *                  line search and a plan whose tiles fit in shared memory next to the staged tables. Measured slower than GRAM_CHOLESKY
 *                  at thousands of instances (three instances per SM cannot hide the sweep's dependent chains; profiles/), so AUTO does
 *                  not pick it. */
typedef enum mb2_fused_mode {
  MB2_FUSED_AUTO = 0,          /* PERSISTENT when the batch is a single wave of instance groups (<= 3 per SM), else GRAM_CHOLESKY when strips / tiles fit in
                                  shared memory, else OFF */
  MB2_FUSED_OFF = 1,
  MB2_FUSED_PERSISTENT = 2,    /* or MB2_ERR_UNSUPPORTED */
  MB2_FUSED_GRAM_CHOLESKY = 3  /* or MB2_ERR_UNSUPPORTED */
} mb2_fused_mode;

/* How the Gauss-Newton step is computed from the Jacobian. CHOLESKY = GaussNewtonSolverT / SubsetGaussNewtonSolverT (normal equations,
 * Eigen::LLT; gauss_newton_solver.cpp:248-251). QR = GaussNewtonSolverQRT (character_solver/gauss_newton_solver_qr.cpp:50-150): an online
 * Householder QR of [sqrt(lambda) I; J] (math/online_householder_qr.cpp), the default solver of pymomentum's solve_ik and of the marker
 * tracker; same step up to rounding without squaring the condition number; its line search is the c1 = 1e-4 / g.delta rule (:116-143). */
typedef enum mb2_linear_solver {
  MB2_LINEAR_SOLVER_CHOLESKY = 0,
  MB2_LINEAR_SOLVER_QR = 1,
  /* TrustRegionQRT (character_solver/trust_region_qr.cpp:52-270; LinearSolverType::TrustRegionQR of pymomentum's solve_ik): QR of the
   * Jacobian, a Newton search for the damping that keeps the step inside the trust region, rho-driven radius, steps with rho <= 0 rejected.
   * regularization, do_line_search and use_block_jtj do not apply (the reference's class takes plain SolverOptions + the radius). */
  MB2_LINEAR_SOLVER_TRUST_REGION_QR = 2
} mb2_linear_solver;

/* solver/solver.h:19-34 SolverOptions + solver/gauss_newton_solver.h:17-59 GaussNewtonSolverOptions,
 * field for field, plus device extensions at the end. */
typedef struct mb2_gauss_newton_options {
  uint64_t min_iterations;        /* SolverOptions::minIterations = 1 */
  uint64_t max_iterations;        /* SolverOptions::maxIterations = 2 */
  float threshold;                /* SolverOptions::threshold = 1.0f */
  int32_t verbose;                /* SolverOptions::verbose */
  float regularization;           /* GaussNewtonSolverBaseOptions::regularization = 0.05f */
  int32_t do_line_search;         /* ::doLineSearch = false */
  int32_t use_block_jtj;          /* ::useBlockJtJ = false (same normal equations either way) */
  uint64_t target_rows_per_chunk; /* ::targetRowsPerChunk = SIZE_MAX (accepted, no effect on results) */
  int32_t subset_line_search;     /* 1 = SubsetGaussNewtonSolverT line search (c1=1e-4, g.delta), subset_gauss_newton_solver.cpp:119-141 */
  int32_t jtj_mode;               /* mb2_jtj_mode */
  int32_t store_error_history;    /* keep per-iteration error per instance (solver.h:90 getErrorHistory) */
  int32_t cholesky_mode;          /* mb2_cholesky_mode */
  int32_t fused_mode;             /* mb2_fused_mode */
  int32_t linear_solver;          /* mb2_linear_solver */
  float trust_region_radius;      /* TrustRegionQROptions::trustRegionRadius_ = 1.0f (trust_region_qr.h:23); MB2_LINEAR_SOLVER_TRUST_REGION_QR only */
} mb2_gauss_newton_options;

typedef struct mb2_character mb2_character;             /* Skeleton + ParameterTransform + ParameterLimits on device */
typedef struct mb2_solver_function mb2_solver_function; /* batch of B SkeletonSolverFunctionT<float> */
typedef struct mb2_solver mb2_solver;                   /* batch of B GaussNewtonSolverT<float> */

const char* mb2_last_error(void);
/* number of usable sm_100 devices (0 => every compute call fails with MB2_ERR_CUDA) */
int mb2_device_count(void);
void mb2_default_gauss_newton_options(mb2_gauss_newton_options* opt);

/* ---- Character: Skeleton (character/skeleton.h:22-77, joint.h:18-76) + ParameterTransform
 * (character/parameter_transform.h:62-184; CSR rows = 7*num_joints) ------------------------------- */
int mb2_character_create(int device, int32_t num_joints, const int32_t* parents /*[J], -1 root*/,
                         const float* translation_offsets /*[J*3]*/, const float* pre_rotations /*[J*4] xyzw*/,
                         int32_t num_model_parameters, const int32_t* transform_outer /*[7J+1]*/,
                         const int32_t* transform_inner /*[nnz]*/, const float* transform_values /*[nnz]*/,
                         const float* transform_offsets /*[7J]*/, mb2_character** out);
/* ParameterLimits consumed by LimitErrorFunctionT (character/parameter_limits.h:129) */
int mb2_character_set_parameter_limits(mb2_character* c, int32_t count, const mb2_parameter_limit* limits);
void mb2_character_destroy(mb2_character* c);

/* ---- SkeletonSolverFunctionT<float> x B (character_solver/skeleton_solver_function.h:21-95) ---- */
int mb2_solver_function_create(const mb2_character* c, int32_t batch, mb2_solver_function** out);
void mb2_solver_function_destroy(mb2_solver_function* f);
int32_t mb2_solver_function_num_parameters(const mb2_solver_function* f);    /* getNumParameters */
int32_t mb2_solver_function_actual_parameters(const mb2_solver_function* f); /* getActualParameters */
int32_t mb2_solver_function_batch(const mb2_solver_function* f);
/* getJacobianBlockSize summed and padded to 8 (solver_function.cpp:33-38) */
int32_t mb2_solver_function_jacobian_rows(const mb2_solver_function* f);
/* row stride (leading dimension) of device Jacobian columns */
int32_t mb2_solver_function_jacobian_stride(const mb2_solver_function* f);

/* addErrorFunction(PositionErrorFunctionT) — position_error_function.h:16-73. Constraint topology
 * (parent, offset, weight) is shared by the batch; targets are per instance. Returns block index. */
int mb2_add_position_error_function(mb2_solver_function* f, float weight, float loss_alpha, float loss_c,
                                    int32_t num_constraints, const int32_t* parents, const float* offsets /*[nc*3]*/,
                                    const float* weights /*[nc]*/, int32_t* out_index);
/* Same, with the constraint OFFSETS per instance as well (the reference builds its error functions per batch element,
 * pymomentum/tensor_ik/tensor_ik.cpp:136-140: offsets and targets may differ from element to element; the parent joints fix the
 * sparsity pattern and stay shared). Per-instance record (mb2_set_targets): [B][nc*6] = target xyz, offset xyz per constraint. */
int mb2_add_position_error_function_instanced(mb2_solver_function* f, float weight, float loss_alpha, float loss_c, int32_t num_constraints,
                                              const int32_t* parents, const float* weights /*[nc]*/, int32_t* out_index);
/* addErrorFunction(PlaneErrorFunctionT) — plane_error_function.h:20-101, .cpp:49-70: signed distance of T_parent * offset to the plane
 * (normal, d), one residual row per constraint; above != 0 is the half-plane mode (only val < 0 is penalised). Per-instance targets
 * (mb2_set_targets): [B][nc*4] = normal xyz (normalised as in PlaneDataT's ctor), d. kLegacyWeight = 1e-4 (.h:83). */
int mb2_add_plane_error_function(mb2_solver_function* f, float weight, float loss_alpha, float loss_c, int32_t above,
                                 int32_t num_constraints, const int32_t* parents, const float* offsets /*[nc*3]*/,
                                 const float* weights /*[nc]*/, int32_t* out_index);
/* addErrorFunction(ModelParametersErrorFunctionT) — model_parameters_error_function.h/.cpp: row sqrt(weight * kMotionWeight) * w_i *
 * (theta_i - target_i) for every enabled parameter with target weight w_i > 0 (kMotionWeight = 1e-1, .h:61). Per-instance targets
 * (mb2_set_targets): [B][numParams] target parameters; target_weights [numParams] is shared by the batch. */
int mb2_add_model_parameters_error_function(mb2_solver_function* f, float weight, const float* target_weights /*[numParams]*/,
                                            int32_t* out_index);
/* addErrorFunction(OrientationErrorFunctionT / OrientationRotDiffErrorFunctionT) —
 * orientation_error_function.h:16-108; offsets are normalised as in OrientationDataT's ctor. */
int mb2_add_orientation_error_function(mb2_solver_function* f, float weight, float loss_alpha, float loss_c,
                                       int32_t rot_diff, int32_t num_constraints, const int32_t* parents,
                                       const float* offsets /*[nc*4]*/, const float* weights /*[nc]*/, int32_t* out_index);
/* addErrorFunction(StateErrorFunctionT) — state_error_function.h:35-117 (setWeights, setTargetWeights) */
int mb2_add_state_error_function(mb2_solver_function* f, float weight, int32_t rotation_error_type, float pos_wgt,
                                 float rot_wgt, const float* target_position_weights /*[J]*/,
                                 const float* target_rotation_weights /*[J]*/, int32_t* out_index);
/* addErrorFunction(LimitErrorFunctionT) over the character's limits — limit_error_function.h:25-119 */
int mb2_add_limit_error_function(mb2_solver_function* f, float weight, float loss_alpha, float loss_c, int32_t* out_index);
/* SkeletonErrorFunctionT::setWeight (skeleton_error_function.h:45-47) */
int mb2_set_error_function_weight(mb2_solver_function* f, int32_t index, float weight);
/* Per-instance targets: Position [B*nc*3] (setConstraints targets), Orientation [B*nc*4] xyzw
 * (normalised on upload), State [B*J*8] = (t, q xyzw, s) (setTargetState). */
int mb2_set_targets(mb2_solver_function* f, int32_t index, const float* targets);
int mb2_set_targets_device(mb2_solver_function* f, int32_t index, const float* targets_device, void* cuda_stream);
/* Optional per-instance constraint weights [B*nc] for a Position/Orientation block (ConstraintData::weight) */
int mb2_set_constraint_weights(mb2_solver_function* f, int32_t index, const float* weights, int32_t per_instance);
/* the same from device memory, [B][nc] contiguous, on `cuda_stream` (NULL = handle stream) */
int mb2_set_constraint_weights_device(mb2_solver_function* f, int32_t index, const float* weights_device, void* cuda_stream);
/* SolverFunctionT::setEnabledParameters(ParameterSet) — skeleton_solver_function.cpp:45-61 */
int mb2_solver_function_set_enabled_parameters(mb2_solver_function* f, const uint64_t bits[MB2_PARAMETER_SET_WORDS]);

/* SolverFunctionT::getError — skeleton_solver_function.cpp:64-83 (value rounded through float). */
int mb2_solver_function_get_error(mb2_solver_function* f, const float* parameters /*[B*n]*/, double* errors /*[B]*/);
/* SolverFunctionT::getJacobian — solver_function.cpp:22-71. jacobian [B][n][rows] (column-major per
 * instance, rows = mb2_solver_function_jacobian_rows), residual [B][rows]. */
int mb2_solver_function_get_jacobian(mb2_solver_function* f, const float* parameters, float* jacobian, float* residual,
                                     double* errors, int32_t* actual_rows);
/* getJacobian with parameters and result on the device (no copy): *jacobian_device points at the handle's Jacobian buffer,
 * [B][n + 1][ld] floats, column c of instance b at ((b * (n + 1)) + c) * ld, column n = residual; rows beyond
 * mb2_solver_function_jacobian_rows are zero. Valid until the next call on the handle. */
int mb2_solver_function_get_jacobian_device(mb2_solver_function* f, const float* parameters_device, const float** jacobian_device, int32_t* ld, void* cuda_stream);
/* SolverFunctionT::getJtJR — solver_function.cpp:74-121. jtj [B][ap][ap] (lower triangle valid,
 * ap = actual parameters), jtr [B][ap]. */
int mb2_solver_function_get_jtjr(mb2_solver_function* f, const float* parameters, int32_t jtj_mode, float* jtj, float* jtr,
                                 double* errors);
/* Skeleton state after initializeJacobianComputation (skeleton_state.cpp:87-121): [B][J][8] (t,q,s) */
int mb2_solver_function_get_skeleton_state(mb2_solver_function* f, const float* parameters, float* state);

/* ---- GaussNewtonSolverT<float> x B (solver/gauss_newton_solver.h:67-137, solver/solver.h:36-100) ---- */
int mb2_solver_create(mb2_solver_function* f, const mb2_gauss_newton_options* opt, mb2_solver** out);
void mb2_solver_destroy(mb2_solver* s);
int mb2_solver_set_options(mb2_solver* s, const mb2_gauss_newton_options* opt); /* setOptions */
/* SolverT::setEnabledParameters — solver.cpp:41-48 (forwards to the solver function) */
int mb2_solver_set_enabled_parameters(mb2_solver* s, const uint64_t bits[MB2_PARAMETER_SET_WORDS]);
/* SolverT::solve for every instance — solver.cpp:50-128. parameters [B*n] in/out (host). errors[b] is
 * the objective before the last update (what solve() returns); iterations[b] = number of
 * doIteration calls; status[b] = mb2_instance_status. Any of errors/iterations/status may be NULL. */
int mb2_solver_solve(mb2_solver* s, float* parameters, double* errors, int32_t* iterations, int32_t* status);
/* The same call in two halves, for callers that keep several handles busy (e.g. the buckets of a mixed-rig batch): solve_async queues
 * the H2D copy, the solve and the D2H copy on the handle's stream and returns (use pinned host memory for a truly asynchronous copy);
 * wait blocks until they are done and returns the per-instance results. */
int mb2_solver_solve_async(mb2_solver* s, float* parameters);
int mb2_solver_wait(mb2_solver* s, double* errors, int32_t* iterations, int32_t* status);
/* Same with parameters resident on the device, on `cuda_stream` (NULL = handle stream). Results are fetched with
 * mb2_solver_get_results after synchronising. The fused single-kernel path (default options on a rig whose tiles fit in shared
 * memory, no line search) is fully asynchronous; the multi-kernel path is asynchronous up to min_iterations and then reads the
 * device-side active counter every fourth iteration (a host round trip on `cuda_stream`) so that a converged batch stops early. */
int mb2_solver_solve_device(mb2_solver* s, float* parameters_device, void* cuda_stream);
int mb2_solver_get_results(mb2_solver* s, double* errors, int32_t* iterations, int32_t* status);
/* getErrorHistory (solver.h:90): [B][max_iterations], valid up to iterations[b] */
int mb2_solver_get_error_history(mb2_solver* s, double* history);
/* sum over the batch of iterations executed / kernels launched by the last solve (for throughput) */
int mb2_solver_get_counters(mb2_solver* s, uint64_t* total_iterations, uint64_t* kernel_launches);
/* device time of the dominant kernels in the last solve, milliseconds (CUDA events on the handle's
 * stream); index: 0 = FK+Jacobian, 1 = JtJ/Jtr, 2 = Cholesky/update, 3 = error-only. Enabled by
 * mb2_solver_set_profiling(s, 1); off by default (events serialise the stream). Level 2 additionally runs the instrumented
 * instantiations of the fused kernels (per-phase SM cycles, mb2_solver_get_fused_profile); those are slower, so take kernel times
 * at level 1 and phase shares at level 2. */
int mb2_solver_set_profiling(mb2_solver* s, int32_t enabled);
int mb2_solver_get_phase_times(mb2_solver* s, double ms[4], uint64_t launches[4]);
/* Per-instance algorithmic sizes of the plan the last solve ran on (roofline accounting in bench.py; no reference
 * counterpart): [0] structurally non-zero Jacobian entries written per iteration, [1] device Jacobian columns (without
 * the residual column), [2] ldJ, [3] parameters in the normal equations (ns), [4] 16x16 tiles held by the tile-sparse
 * Cholesky (0: dense Eigen-structured kernel), [5] tile multiply-accumulate blocks per factorisation, [6] levels of the
 * tile elimination tree, [7] residual rows m (row groups aligned to 4 in the strip layout), [8] floats per instance of the Jacobian in
 * strip layout (0: K-major matrix), [9] multiply-accumulates per instance of the tile-sparse Gram kernel, [10] its strip pairs,
 * [11] instance groups per CTA of the fused persistent kernel (0: the plan does not fit / was not built for it). */
int mb2_solver_get_plan_stats(mb2_solver* s, int64_t stats[12]);
/* Which fused kernel the last solve ran (fused: 0 none, 1 persistent whole-solve kernel, 2 Gram + Cholesky per iteration) and, after
 * mb2_solver_set_profiling(s, 1), its device time and the per-phase SM cycles of one instance group (persistent kernel: over its whole
 * share of the batch; Gram + Cholesky: CTA 0 summed over the iterations, [0] = prologue, [1..4] unused): [0] work fetch + theta load, [1] ParameterTransform
 * + strip zero fill, [2] FK sweep, [3] residual/units, [4] Jacobian cells, [5] Gram (J^T J, J^T r), [6] tiles out of TMEM,
 * [7] Cholesky diagonal tiles, [8] panel tiles, [9] updates, [10] backward substitution, [11] update + bookkeeping + write-back.
 * groups = instance groups per CTA of the persistent kernel. Any output pointer may be NULL. No reference counterpart. */
int mb2_solver_get_fused_profile(mb2_solver* s, int32_t* fused, int32_t* groups, double* kernel_ms, uint64_t phase_cycles[12]);

/* ---- Mixed-rig batches (BASELINE.json configs[4]): heterogeneous instances -> buckets sharing one plan -> batched solves --------
 * The reference treats a batch element by element (pymomentum/tensor_ik/tensor_ik.cpp:127-177 builds error functions, solver function
 * and solver per element). Here instances are bucketed by (rig, constraint parents): inside a bucket offsets, targets, weights and
 * the number of constraints are per instance (shorter instances are padded with zero-weight constraints, at most granule - 1 of them),
 * every bucket is one batched solve, results come back in input order. Position constraints (+ optionally the rig's ParameterLimits). */
typedef struct mb2_mixed_batch mb2_mixed_batch;
const char* mb2_mixed_batch_last_error(void);
int mb2_mixed_batch_create(int device, int32_t granule /*constraints; <= 0: 8*/, mb2_mixed_batch** out);
void mb2_mixed_batch_destroy(mb2_mixed_batch* b);
/* the character must outlive the batch (ownership as everywhere in momentum: non-owning references) */
int mb2_mixed_batch_add_rig(mb2_mixed_batch* b, const mb2_character* c, int32_t num_parameters, int32_t* rig_id);
int mb2_mixed_batch_use_limits(mb2_mixed_batch* b, int32_t enabled, float weight); /* LimitErrorFunctionT over each rig's limits */
int mb2_mixed_batch_add_instance(mb2_mixed_batch* b, int32_t rig_id, int32_t num_constraints, const int32_t* parents, const float* offsets /*[nc*3]*/,
                                 const float* weights /*[nc]*/, const float* targets /*[nc*3]*/, const float* theta0 /*[n of the rig]*/, int32_t* instance_id);
int mb2_mixed_batch_set_parameters(mb2_mixed_batch* b, int32_t instance_id, const float* theta0);
int mb2_mixed_batch_solve(mb2_mixed_batch* b, const mb2_gauss_newton_options* opt);
int mb2_mixed_batch_get_result(mb2_mixed_batch* b, int32_t instance_id, float* theta, double* error, int32_t* iterations, int32_t* status);
int mb2_mixed_batch_get_results(mb2_mixed_batch* b, float* theta, const int64_t* theta_offsets, double* errors, int32_t* iterations, int32_t* status);
/* [0] instances, [1] buckets, [2] residual rows before padding, [3] after, [4] largest bucket, [5] singleton buckets */
int mb2_mixed_batch_stats(mb2_mixed_batch* b, int64_t stats[6]);
/* [0] rig, [1] instances, [2] constraints (padded), [3] Gauss-Newton iterations of the last solve */
int mb2_mixed_batch_bucket_info(mb2_mixed_batch* b, int32_t bucket, int64_t info[4]);

/* ---- single-process multi-GPU: one batch sharded over several devices (north_star "host code stays C++") ----
 * The reference solves a batch element by element on host threads (pymomentum/tensor_ik/tensor_ik.cpp:127-177, dispenso::parallel_for);
 * instances are independent, so a batch of total_batch instances is cut into contiguous blocks, one per device, each block an ordinary
 * mb2_solver on its own device driven by its own host thread. There is no data-path collective: the only cross-device quantity is the
 * aggregate {sum of final errors, total iterations, instances that ended with status OK}, summed on the host from the per-instance
 * results (the 24-byte all-reduce of SURVEY 8e; with one process there is nothing to send over NVLink).
 *
 * Replicas: mb2_character_clone puts the rig on another device; mb2_solver_function_clone copies the DEFINITION of a solver function
 * (error-function blocks, shared constraint weights, block weights, enabled parameters - no targets) for a new batch size. */
int mb2_character_clone(const mb2_character* c, int device, mb2_character** out);
int mb2_solver_function_clone(const mb2_solver_function* f, const mb2_character* c_on_device, int32_t batch, mb2_solver_function** out);
int mb2_character_device(const mb2_character* c);
const mb2_character* mb2_solver_function_character(const mb2_solver_function* f);
int32_t mb2_solver_function_num_error_functions(const mb2_solver_function* f);
int32_t mb2_solver_function_target_size(const mb2_solver_function* f, int32_t index); /* floats per instance of block `index` */

typedef struct mb2_sharded_solver mb2_sharded_solver;
const char* mb2_sharded_last_error(void);
/* `prototype` defines the problem (it is only read, on its own device, and may be destroyed afterwards); devices[] may name a device more
 * than once (several shards on one GPU). Shard k holds instances [first_k, first_k + count_k), count = total_batch / num_devices rounded
 * so that the counts differ by at most one. */
int mb2_sharded_solver_create(const mb2_solver_function* prototype, int32_t total_batch, int32_t num_devices, const int32_t* devices,
                              const mb2_gauss_newton_options* opt, mb2_sharded_solver** out);
void mb2_sharded_solver_destroy(mb2_sharded_solver* s);
int32_t mb2_sharded_solver_num_shards(const mb2_sharded_solver* s);
int mb2_sharded_solver_shard_info(const mb2_sharded_solver* s, int32_t shard, int32_t info[3] /* device, first instance, count */);
int mb2_sharded_solver_set_options(mb2_sharded_solver* s, const mb2_gauss_newton_options* opt);
/* targets[total_batch][size of block index] in instance order, host memory (SkeletonErrorFunction::setConstraints per instance) */
int mb2_sharded_solver_set_targets(mb2_sharded_solver* s, int32_t index, const float* targets);
/* SolverT::solve over the whole batch: parameters[total_batch * n] in/out (host), per-instance results optional */
int mb2_sharded_solver_solve(mb2_sharded_solver* s, float* parameters, double* errors, int32_t* iterations, int32_t* status);
/* aggregate of the last solve: [0] sum of final errors, [1] total Gauss-Newton iterations, [2] instances with status MB2_INSTANCE_OK */
int mb2_sharded_solver_get_aggregate(const mb2_sharded_solver* s, double aggregate[3]);

#ifdef __cplusplus
}
#endif
#endif /* MOMENTUM_B200_H_ */

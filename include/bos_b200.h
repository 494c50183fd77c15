/* bos_b200.h -- C ABI of the B200-native Gauss-Newton solver for 2D bearing-only SLAM.
 *
 * This is the drop-in boundary under the reference's C++ class surface
 * (torchipeppo/prb-project-bearing-only-slam).  The reference has no FFI of its own: the
 * entry points below are what a binding for `proj02::Solver` / `triangulate_landmarks`
 * would call, and each cites the reference interface it replaces.  Plain pointers and
 * sizes only; every call returns an int status (0 = ok); no exception crosses the
 * boundary; a context owns its device memory and stream and is used from one host
 * thread at a time.  There is NO CPU fallback: without a CUDA device every compute call
 * fails with BOS_ERR_CUDA.
 *
 * Index conventions (bit-exact with the reference):
 *   pose stix      = insertion order of State::add_pose            (framework/state.cpp:20-30)
 *   landmark stix  = insertion order of State::add_landmark        (framework/state.cpp:32-41);
 *                    after triangulate_landmarks that is ASCENDING landmark id
 *                    (slam/triangulation.cpp:65-74)
 *   delta layout   = [3*NP pose | 2*NL landmark]                   (framework/state.cpp:69-80)
 *   pose on the wire = (x, y, c, s) with c = R(0,0), s = R(1,0) of the reference's
 *                    Isometry2f (never re-orthonormalised, as in the reference)
 * All floating point crosses the boundary as double whatever the compute precision.
 */
#ifndef BOS_B200_H
#define BOS_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define BOS_API __attribute__((visibility("default")))
#else
#define BOS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bos_ctx bos_ctx;
typedef struct bos_batch bos_batch;

enum {
    BOS_OK = 0,
    BOS_ERR_INVALID = 1,   /* bad argument (null pointer, index out of range, self-loop ...) */
    BOS_ERR_CUDA = 2,      /* CUDA runtime failure, or no device */
    BOS_ERR_STATE = 3,     /* call order (e.g. step before upload_problem) */
    BOS_ERR_NCCL = 4,      /* NCCL missing or failed */
    BOS_ERR_NOMEM = 5
};

enum { BOS_PRECISION_F64 = 0, BOS_PRECISION_F32 = 1 };
/* BOS_SOLVER_SPARSE_CHOLESKY: Schur complement + Cholesky of the reduced pose system in SKYLINE storage: the symbolic phase (the
 * envelope of the pose-only Schur complement in pose order: the analogue of SimplicialLDLT::analyzePattern, slam/solver.cpp:77-80) runs once
 * per uploaded pattern, the numeric phase runs the dense solver's blocked DMMA kernels on the windows the envelope allows. */
enum { BOS_SOLVER_AUTO = 0, BOS_SOLVER_DENSE_CHOLESKY = 1, BOS_SOLVER_PCG = 2, BOS_SOLVER_SPARSE_CHOLESKY = 3 };

typedef struct bos_options {
    int device;               /* CUDA ordinal */
    int precision;            /* BOS_PRECISION_* ; F64 is the 1e-9 parity path, F32 is the reference's own precision */
    int solver;               /* BOS_SOLVER_* ; AUTO = dense Cholesky when 3*NP <= dense_max_dim, else PCG */
    int dense_max_dim;        /* default 192: measured on B200 (profiles/solver_sweep_r02.jsonl) the persistent PCG beats the dense
                                 Cholesky from 3*NP ~ 300 on (0.6 vs 0.8 ms) and by 44x at 3*NP = 18000; the dense path stays available through
                                 BOS_SOLVER_DENSE_CHOLESKY (BASELINE config 3 names it) */
    double kernel_threshold;  /* slam/solver.cpp:16  default 1.0  */
    double damping;           /* slam/solver.cpp:17  default 0.01 */
    int pcg_max_iters;        /* default 20000 (block-Jacobi needs ~5300 CG iterations on a 200k-pose odometry chain) */
    double pcg_rtol;          /* stop when sqrt(r^T M^-1 r) <= rtol * its initial value; default 1e-10 */
    int pcg_variant;          /* 0 = one persistent cooperative kernel for the whole PCG solve (default; falls back to 1 when a CTA's
                                 chunk of poses does not fit: more than 2048 poses per chunk or more dynamic shared memory than an SM
                                 has), 1 = classic loop of small kernels */
    int pcg_precond;          /* fused kernel, FP64 only: 0 = chain blocks + coarse space (default): per chunk of ~NP/148 poses the
                                 block-tridiagonal matrix of the Schur diagonal blocks and the pose-pose blocks of consecutive poses,
                                 factorised once per solve and applied exactly (block-Jacobi with chunk-sized blocks), plus a
                                 Galerkin coarse space of piecewise-linear hats over the chunks (~70x fewer CG iterations than 3x3
                                 blocks on odometry chains); 2 = the chain blocks alone; 1 = the 3x3 block-Jacobi preconditioner
                                 (also what FP32 and pcg_variant 1 run).  All converge to the same solution at pcg_rtol. */
    int pcg_coarse_nodes;     /* coarse-space nodes per chunk of the default preconditioner: 0 = default (4), 1 = the round-1 layout (one
                                 hat per chunk); rounded so that segments are whole groups of 32 poses, at most 8 */
    int pcg_coarse_refresh;   /* the coarse operator's inverse is kept across GN steps.  1 = rebuild for every solve; otherwise (0 = default = 8) it is
                                 rebuilt when that pays -- once the CG iterations spent above the post-rebuild count add up to the measured price
                                 of a rebuild (about 31 iterations on the 200 k-pose world) --, at the latest after 8 x this many solves, and
                                 whenever the state is replaced through the API or the CG iteration count drifts up by 25 %.  Any SPD coarse
                                 operator is a valid preconditioner: only the iteration count depends on it, never the solution */
    int reserved[4];
} bos_options;

/* Per-iteration outputs.  The reference prints none of these; chi2 is defined as the sum of the
 * per-edge error_omeganorm it already computes BEFORE the robust scaling (slam/solver.cpp:38,55). */
typedef struct bos_stats {
    double chi2_bearing;
    double chi2_odometry;
    int64_t over_bearing;     /* edges whose error_omeganorm exceeded kernel_threshold */
    int64_t over_odometry;
    double delta_inf;         /* max |dx| */
    int solver_status;        /* 0 ok, 1 = non-positive pivot / CG breakdown (the reference's "not SPD" console message,
                                 solver.cpp:82-84), 2 = the PCG stopped at pcg_max_iters before reaching pcg_rtol */
    int solver_used;          /* BOS_SOLVER_DENSE_CHOLESKY, BOS_SOLVER_PCG or BOS_SOLVER_SPARSE_CHOLESKY */
    int pcg_iterations;
    int gpu_launches;         /* kernels launched by the last step() */
    float ms_linearize;       /* CUDA-event times of the last step(), on the context's stream */
    float ms_solve;
    float ms_update;
    float ms_allreduce;
    int precond_used;         /* PCG: the preconditioner that actually ran, in bos_options.pcg_precond numbering (a request for the
                                 chain / coarse-space preconditioner is downgraded when its factors do not fit in shared memory, in
                                 FP32, and after a CG breakdown); -1 when the dense solver ran */
    int pcg_resolves;         /* 1: the PCG broke down under the chain / coarse preconditioner and was run again with 3x3 blocks */
    double state_digest;      /* sum of every scalar of the state AFTER the update (x + y + c + s per pose, x + y per landmark):
                                 lets two runs (1 GPU vs N GPUs, device vs host stepping) be compared without downloading the state */
} bos_stats;

BOS_API void bos_default_options(bos_options* o);
BOS_API int bos_version(void);

/* Solver::Solver (slam/solver.cpp:5-18): create a context; problem and state are uploaded separately. */
BOS_API int bos_create(const bos_options* opts, bos_ctx** out);
BOS_API int bos_destroy(bos_ctx* ctx);
BOS_API const char* bos_last_error(const bos_ctx* ctx);

/* Solver::set_kernel_threshold / set_damping_factor (slam/solver.cpp:20-25). */
BOS_API int bos_set_kernel_threshold(bos_ctx* ctx, double kt);
/* Robust kernel flavour (opt-in; the default is the reference's).  BOS_ROBUST_REFERENCE: an edge whose chi2 = e^T Omega e exceeds
 * kernel_threshold has its ERROR scaled by w = sqrt(kt / chi2), J and Omega untouched (slam/solver.cpp:37-41, 54-58).  BOS_ROBUST_IRLS: the
 * standard iteratively re-weighted form of the same (Huber-type) weight: Omega <- w Omega, so b is what the reference computes and
 * H += w J^T Omega J instead of J^T Omega J.  IRLS solves run on the stored blocks (dense Cholesky or the classic PCG loop). */
#define BOS_ROBUST_REFERENCE 0
#define BOS_ROBUST_IRLS 1
BOS_API int bos_set_robust_mode(bos_ctx* ctx, int mode);
BOS_API int bos_set_damping_factor(bos_ctx* ctx, double df);

/* The edge vectors + construct_the_permutation (slam/solver.cpp:5-18, 99-125).  Edges arrive with
 * ids already resolved to stix (the reference resolves them per edge per iteration through
 * std::map::at, framework/state.cpp:43-63).  Host buffers are borrowed for the call and copied.
 * b_omega may be NULL (= 1, framework/observation.hpp:17).  o_z is [Eo][3]; o_omega is [Eo][9]
 * row-major and must be symmetric, as utils/g2o_utils.cpp:91-106 builds it (the kernels keep the upper
 * triangle; a non-symmetric matrix is refused with BOS_ERR_INVALID instead of silently differing from
 * the reference's full 3x3 product).
 * Builds the CSR-of-blocks sparsity pattern and the per-edge block slots. */
BOS_API int bos_upload_problem(bos_ctx* ctx, int NP, int NL, int fixed_pose_stix,
                       int64_t Eb, const int32_t* b_pose, const int32_t* b_lm, const double* b_z, const double* b_omega,
                       int64_t Eo, const int32_t* o_src, const int32_t* o_dst, const double* o_z, const double* o_omega);

/* Solver::state (slam/solver.hpp:26). poses [NP][4] = x,y,c,s ; lms [NL][2].  Either may be NULL. */
BOS_API int bos_set_state(bos_ctx* ctx, const double* poses_xycs, const double* lms_xy);
BOS_API int bos_get_state(bos_ctx* ctx, double* poses_xycs, double* lms_xy);

/* The three phases of Solver::step (slam/solver.cpp:27-97), callable one by one for parity tests:
 *   linearize : lines 28-69  (errors, Jacobians, robust kernel, H/b accumulation, damping)
 *   solve     : lines 72-94  (gauge fix, factorise, solve, re-expand dx)
 *   update    : line  96     (State::apply_boxplus, framework/state.cpp:69-80) */
BOS_API int bos_linearize(bos_ctx* ctx);
BOS_API int bos_solve(bos_ctx* ctx);
BOS_API int bos_update(bos_ctx* ctx);   /* needs an increment from bos_solve / bos_upload_delta; applies it once (BOS_ERR_STATE otherwise) */

/* Solver::step(): exactly one GN iteration including the state update, state resident on the device. */
BOS_API int bos_step(bos_ctx* ctx, bos_stats* stats);
/* Same, through HOST buffers: uploads the state, steps, downloads the new state (the reference's
 * callers read solver.state after every step, executables/bearing_only_slam.cpp:31-36). */
BOS_API int bos_step_host(bos_ctx* ctx, double* poses_xycs_inout, double* lms_xy_inout, bos_stats* stats);
BOS_API int bos_get_stats(bos_ctx* ctx, bos_stats* stats);

/* Extension beyond the reference (the reference has a FIXED damping and never rejects a step, slam/solver.cpp:64-69):
 * one Levenberg-Marquardt iteration.  Takes a GN step with the current damping, relinearizes, and compares the total chi2
 * (chi2_bearing + chi2_odometry, the pre-kernel error_omeganorm sums) before and after: if it did not decrease the state is
 * restored and the damping multiplied by 10, otherwise the step is kept and the damping divided by 3 (clamped to [1e-9, 1e9]).
 * stats describes the GN step taken; chi2_after / accepted / damping_next may be NULL.  Opt-in: bos_step never does this. */
BOS_API int bos_step_lm(bos_ctx* ctx, bos_stats* stats, double* chi2_after, int* accepted, double* damping_next);

/* triangulate_landmarks (slam/triangulation.cpp:5-74) on the uploaded bearing edges and the current
 * poses; writes all NL landmarks of the device state.  single_obs_count (may be NULL) receives the
 * number of landmarks with exactly one observation (the reference's console warning, :38-42). */
BOS_API int bos_triangulate(bos_ctx* ctx, int* single_obs_count);

/* ---- context-free forms (what the reference's free function / per-edge methods would bind) -------------------- */
/* triangulate_landmarks(State&, const BearingObservationVector&) (slam/triangulation.hpp:8, triangulation.cpp:5-74):
 * poses_xycs [NP][4]; bearing edges with pose stix and landmark index 0..NL-1 in ASCENDING landmark id (the stix
 * order in which the reference adds them, triangulation.cpp:65-74); writes lms_xy [NL][2].  opts may be NULL. */
BOS_API int bos_triangulate_landmarks(const bos_options* opts, int NP, const double* poses_xycs, int64_t Eb, const int32_t* b_pose,
                              const int32_t* b_lm, const double* b_z, int NL, double* lms_xy, int* single_obs_count);
/* Solver::error_and_jacobian (slam/solver.hpp:35-37) for n independent edges, evaluated on the device:
 * bearing: poses_xycs [n][4], lms_xy [n][2], z [n] -> err [n], jac [n][5] = [J_pose | J_lm];
 * odometry: src/dst [n][4], z [n][3] -> err [n][3], jac [n][18] = 3x6 row-major [J_src | J_dst]. */
BOS_API int bos_eval_bearing_edges(const bos_options* opts, int64_t n, const double* poses_xycs, const double* lms_xy, const double* z,
                           double* err, double* jac5);
BOS_API int bos_eval_odometry_edges(const bos_options* opts, int64_t n, const double* src_xycs, const double* dst_xycs, const double* z3,
                            double* err3, double* jac18);

/* ---- parity / inspection (tests and the harness; not on the hot path) ------------------------ */
typedef struct bos_pattern_info {
    int64_t n_hpl;        /* unique (pose, landmark) blocks, 3x2 */
    int64_t n_hpp_off;    /* unique pose-pose off-diagonal blocks, 3x3, stored as H[lo][hi] */
    int64_t csc_n;        /* N - 3 */
    int64_t csc_nnz;      /* scalar nnz of H_nofixed, both triangles (what SimplicialLDLT is handed) */
    int64_t N;
    int64_t vals_len;     /* scalars in the reducible value buffer */
} bos_pattern_info;
BOS_API int bos_pattern_info_get(bos_ctx* ctx, bos_pattern_info* out);
/* Block coordinates, sorted: hpl_pose/hpl_lm [n_hpl]; off_lo/off_hi [n_hpp_off]; per-edge slots. */
BOS_API int bos_download_pattern(bos_ctx* ctx, int32_t* hpl_pose, int32_t* hpl_lm, int32_t* off_lo, int32_t* off_hi,
                         int64_t* b_slot, int64_t* o_slot);
/* Block values after linearize: Hpp [NP][9], Hll [NL][4], Hpl [n_hpl][6], Hoff [n_hpp_off][9], b [N]. Any may be NULL. */
BOS_API int bos_download_blocks(bos_ctx* ctx, double* Hpp, double* Hll, double* Hpl, double* Hoff, double* b);
/* H_nofixed as scalar CSC (sorted rows, both triangles) and b_nofixed: slam/solver.cpp:72-75. */
BOS_API int bos_download_csc(bos_ctx* ctx, int32_t* colptr, int32_t* rowidx, double* val, double* b_nofixed);
BOS_API int bos_download_delta(bos_ctx* ctx, double* delta);
BOS_API int bos_upload_delta(bos_ctx* ctx, const double* delta);
/* Per-edge error (before the robust scaling) and Jacobian in the caller's edge order:
 * err_b [Eb], jac_b [Eb][5] = [J_pose | J_lm]; err_o [Eo][3], jac_o [Eo][18] = 3x6 row-major [J_src | J_dst]. */
BOS_API int bos_edge_terms(bos_ctx* ctx, double* err_b, double* jac_b, double* err_o, double* jac_o);

/* Host-only view of the pattern builder (integer work, no device needed): what bos_upload_problem builds,
 * exposed so the sparsity pattern, edge-to-block indexing and sharding can be checked bit-exactly anywhere. */
typedef struct bos_host_pattern bos_host_pattern;
BOS_API int bos_host_pattern_create(int NP, int NL, int fixed_pose_stix, int64_t Eb, const int32_t* b_pose, const int32_t* b_lm,
                            int64_t Eo, const int32_t* o_src, const int32_t* o_dst, bos_host_pattern** out);
BOS_API int bos_host_pattern_destroy(bos_host_pattern* p);
BOS_API int bos_host_pattern_info(const bos_host_pattern* p, bos_pattern_info* out);
BOS_API int bos_host_pattern_get(const bos_host_pattern* p, int32_t* hpl_pose, int32_t* hpl_lm, int32_t* off_lo, int32_t* off_hi,
                         int64_t* b_slot, int64_t* o_slot, int32_t* csc_colptr, int32_t* csc_rowidx);
/* 64-bit FNV-1a digest of EVERY integer table of the pattern (block slots, ELL / chunk / tile layouts, adjacency ...): lets a test
 * assert that two builds (e.g. serial and threaded, BOS_PATTERN_THREADS=1) produced identical device layouts. */
BOS_API int bos_host_pattern_checksum(const bos_host_pattern* p, uint64_t* out);
/* On-device problem setup (the reference resolves ids through std::map for every edge of every iteration, framework/state.cpp:43-63,
 * and this library builds its tables on host threads for small problems).
 * bos_set_device_setup(ctx, on): 1 = always, 0 = never, -1 = from 200 000 bearing edges on (the default): bos_upload_problem builds the bearing-edge core of the pattern -- the (pose, landmark)-sorted edge order, the
 *   pose-landmark block slots, the CSR-of-blocks row pointers, the landmark-major slot order and the triangulation rows -- with GPU radix sorts,
 *   scans and run-length kernels; the remaining layouts (tiles, PCG chunks) are derived from them on the host.  The tables are bit-identical to
 *   the host builder's: bos_pattern_checksum equals bos_host_pattern_checksum.  bos_last_setup_ms reports the two parts of the last upload (device part 0 when the host built everything).
 * bos_device_resolve_ids: id -> stix for all edge end points on the device (map::at semantics: an unknown pose id is BOS_ERR_INVALID, a duplicated
 *   pose id resolves to its last insertion; landmark stix = rank of the id among the observed landmark ids, slam/triangulation.cpp:68-73);
 *   lm_ids (capacity Eb) receives the ascending landmark id table, *NL_out its length. */
BOS_API int bos_set_device_setup(bos_ctx* ctx, int on);
BOS_API int bos_last_setup_ms(const bos_ctx* ctx, double* device_core_ms, double* host_ms);
BOS_API int bos_pattern_checksum(bos_ctx* ctx, uint64_t* out);
BOS_API int bos_device_resolve_ids(int device, int NP, const int32_t* pose_ids, int64_t Eb, const int32_t* b_pose_id, const int32_t* b_lm_id, int64_t Eo,
                                   const int32_t* o_src_id, const int32_t* o_dst_id, int32_t* b_pose, int32_t* b_lm, int32_t* o_src, int32_t* o_dst,
                                   int32_t* lm_ids, int32_t* NL_out);
/* Symbolic phase of BOS_SOLVER_SPARSE_CHOLESKY (what SimplicialLDLT::analyzePattern is to the reference, slam/solver.cpp:77-80): the row limit
 * of every 64-column panel of the reduced pose system's skyline (panel_end: n_panels = ceil(3 NP / 64) entries, may be NULL), the rows stored
 * per column and the stored fraction of the lower triangle.  Host only. */
BOS_API int bos_host_pattern_skyline(const bos_host_pattern* p, int32_t* panel_end, int32_t* n_panels, int32_t* rows_per_column, double* fill);
/* The contiguous edge ranges rank `rank` of `nranks` linearizes: out4 = b_begin, b_end, o_begin, o_end. */
BOS_API int bos_host_edge_shard(int64_t Eb, int64_t Eo, int rank, int nranks, int64_t* out4);

/* ---- multi-GPU: edge-sharded linearization, partials combined by an NCCL allreduce ------------ */
#define BOS_NCCL_UID_BYTES 128
BOS_API int bos_nccl_unique_id(char* uid128);
BOS_API int bos_comm_init(bos_ctx* ctx, int rank, int nranks, const char* uid128);
/* reduce_mode 0: allreduce the whole value buffer (H, b); 1: allreduce only the blocks that can overlap between ranks
 * (b, diagonal blocks, pose-pose blocks) and allgather the rank-owned pose-landmark blocks; 2: allreduce those blocks only and
 * leave the pose-landmark blocks rank-local -- enough for the fused PCG solve, which applies that part of the operator from
 * per-edge factors (bos_download_blocks / bos_download_csc then see this rank's pose-landmark blocks only); 3: ownership-based
 * combine: a pose's diagonal block and rhs are complete on the rank whose edge tiles own the pose (the tile its bearing run starts in),
 * every rank computes every pose-pose block itself, so only the landmark blocks and b_l are summed (5 scalars per landmark) and the
 * owned pose ranges are gathered (grouped broadcasts); pose-landmark blocks stay rank-local as in mode 2;
 * 4: mode 3's ownership with NO collective: the bearing kernel itself stores every owned pose block and adds every landmark part straight into
 * EVERY rank's replica through NVLink peer mappings (bos_peer_open), between two cross-GPU barriers -- the combine rides on the build;
 * 5: mode 3's local build, then the combine as bulk PULLS over the same peer mappings: after a cross-GPU barrier one kernel copies the owners'
 * pose ranges from their replicas and sums the landmark parts of all ranks in rank order (every rank ends up with bit-identical H, b), a
 * second barrier, a commit kernel.  No NCCL call in the build. */
BOS_API int bos_set_reduce_mode(bos_ctx* ctx, int reduce_mode);
/* reduce_mode 4 plumbing (ranks = processes of ONE NVSwitch box, at most 8).  After bos_upload_problem and bos_comm_init / bos_set_edge_shard:
 * every rank exports the CUDA IPC handle of its value buffer (+ byte offset inside the underlying allocation), the host program exchanges them
 * (any transport: the Python wrapper uses torch.distributed.all_gather_object) and every rank opens all of them: handles = nranks x
 * BOS_IPC_HANDLE_BYTES, offsets = nranks entries, both indexed by rank (the own entry is ignored).  Collective: from then on every rank must
 * call bos_linearize / bos_step the same number of times.  A rank that never arrives makes the barrier time out: BOS_ERR_NCCL from bos_get_stats. */
#define BOS_IPC_HANDLE_BYTES 64
BOS_API int bos_peer_export(bos_ctx* ctx, void* handle64, int64_t* offset);
BOS_API int bos_peer_open(bos_ctx* ctx, const void* handles, const int64_t* offsets);
/* Without NCCL: shard bookkeeping only (used by the host-side tests): this rank linearizes its
 * contiguous range of the pose-sorted edges. */
BOS_API int bos_set_edge_shard(bos_ctx* ctx, int rank, int nranks);
BOS_API int bos_get_edge_shard(bos_ctx* ctx, int64_t* b_begin, int64_t* b_end, int64_t* o_begin, int64_t* o_end);

/* ---- batched small problems: one GN iteration per problem per launch ------------------------- */
/* nprob problems sharing one topology (stix arrays, omegas, fixed pose); measurements and states are
 * per problem: b_z [nprob][Eb], o_z [nprob][Eo][3]. */
BOS_API int bos_batch_create(const bos_options* opts, int nprob, int NP, int NL, int fixed_pose_stix,
                     int Eb, const int32_t* b_pose, const int32_t* b_lm, const double* b_z, const double* b_omega,
                     int Eo, const int32_t* o_src, const int32_t* o_dst, const double* o_z, const double* o_omega,
                     bos_batch** out);
BOS_API int bos_batch_destroy(bos_batch* b);
BOS_API int bos_batch_set_states(bos_batch* b, const double* poses_xycs, const double* lms_xy); /* [nprob][NP][4], [nprob][NL][2] */
BOS_API int bos_batch_get_states(bos_batch* b, double* poses_xycs, double* lms_xy);
/* chi2 [nprob][2] (bearing, odometry), delta_inf [nprob], status [nprob]; any may be NULL. */
BOS_API int bos_batch_step(bos_batch* b, double* chi2, double* delta_inf, int32_t* status);
/* Steps resident on the device without reading results back; elapsed_ms (may be NULL) is the
 * CUDA-event time of the n_steps launches. */
BOS_API int bos_batch_step_device(bos_batch* b, int n_steps, float* elapsed_ms);
BOS_API const char* bos_batch_last_error(const bos_batch* b);

#ifdef __cplusplus
}
#endif
#endif /* BOS_B200_H */

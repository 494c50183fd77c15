// solver.hpp -- proj02::Solver with the reference's class surface (slam/solver.hpp:21-92); the body of step() is the
// CUDA path behind include/bos_b200.h.  One Solver owns one bos_ctx (one GPU); use it from one thread at a time.
#pragma once

#include <string>

#include "../../../include/bos_b200.h"
#include "../framework/observation.hpp"
#include "../framework/state.hpp"

namespace proj02 {

// extension: what the reference fixes at compile time (WHICH_SOLVER, slam/solver.hpp:14-17) or does not have
struct SolverOptions {
    int device = 0;
    bool fp32 = false;                 // false: FP64 arithmetic (parity path); true: the reference's own precision
    int solver = BOS_SOLVER_AUTO;      // Schur + dense Cholesky for small problems, Schur + preconditioned CG for large ones
    int dense_max_dim = 192;     // measured crossover on B200: the PCG wins from 3 NP ~ 300 on
    int pcg_max_iters = 20000;
    double pcg_rtol = 1e-10;
    int pcg_variant = 0;               // 0 persistent cooperative kernel, 1 classic multi-kernel loop
    int pcg_precond = 0;               // 0 block-tridiagonal chain preconditioner, 1 3x3 block-Jacobi (see bos_b200.h)
};

class Solver {
 public:
    // as in the reference: the solver holds COPIES of the state and of both edge vectors (slam/solver.cpp:5-9);
    // callers read `state` after every step().
    State state;
    BearingObservationVector bearing_observations;
    OdometryObservationVector odometry_observations;

    Solver(const State& state, const BearingObservationVector& bear_obs, const OdometryObservationVector& odom_obs, const int& fixed_pose_id);
    Solver(const State& state, const BearingObservationVector& bear_obs, const OdometryObservationVector& odom_obs, const int& fixed_pose_id,
           const SolverOptions& options);
    ~Solver();
    Solver(const Solver&) = delete;
    Solver& operator=(const Solver&) = delete;

    void set_kernel_threshold(float kt);   // default 1.0   (slam/solver.cpp:16)
    void set_damping_factor(float df);     // default 0.01  (slam/solver.cpp:17)

    // exactly one Gauss-Newton iteration including the state update (slam/solver.cpp:27-97): linearize + assemble,
    // damping, gauge fix, Schur solve, boxplus -- all on the GPU; `state` is refreshed from the device afterwards.
    // If `state` was modified by the caller since the last step it is uploaded first.
    void step();

    // per-edge evaluation (slam/solver.hpp:35-43).  The analytic pair runs on the device through the C ABI; the
    // numeric Jacobians (central differences, eps = 1e-3, through boxplus) are the reference's own validation aid and
    // stay host-side float code, as in its tests/solver_stuff.cpp.
    void error_and_jacobian(const State& state, const BearingObservation& obs, float& error, SparseMatrixXf& jacobian);
    void error_and_jacobian(const State& state, const OdometryObservation& obs, EPose& error, SparseMatrixXf& jacobian);
    void error_and_numerical_jacobian(const State& state, const BearingObservation& obs, float& error, SparseMatrixXf& jacobian);
    void error_and_numerical_jacobian(const State& state, const OdometryObservation& obs, EPose& error, SparseMatrixXf& jacobian);
    float predict_bearing(const NEPose& pose, const LMPos& lm);
    EPose predict_odometry(const NEPose& src, const NEPose& dst);
    float normalized_angle(float angle);

    // ---- extensions -------------------------------------------------------------------------------------------------
    void step(int iterations, bool mirror_every_step = false);   // several iterations with the state resident on the device
    // opt-in Levenberg-Marquardt iteration (not in the reference: fixed damping, no step rejection): a GN step that is undone,
    // with the damping raised, when the total chi2 does not decrease; returns whether the step was kept
    bool step_lm();
    // opt-in IRLS flavour of the threshold robust kernel (not in the reference, which scales the error only, slam/solver.cpp:37-41): the
    // weight sqrt(kt / chi2) of an over-threshold edge scales its Omega, i.e. H as well as b
    void set_irls(bool on);
    const bos_stats& last_stats() const { return stats_; }       // chi2 (pre-kernel error_omeganorm sums), |dx|_inf, timings
    bool last_step_not_spd() const { return stats_.solver_status != 0; }
    bos_ctx* context() { return ctx_; }

 private:
    void upload_problem();
    void upload_state();
    void download_state();
    void check(int rc, const char* what) const;

    bos_ctx* ctx_ = nullptr;
    bos_options opt_;
    bos_stats stats_;
    int fixed_pose_id_;
    int N_;
    unsigned long long synced_version_ = 0;
    bool device_state_valid_ = false;
};

}  // namespace proj02

// solver.cpp -- host side of proj02::Solver: id -> stix resolution (once, instead of per edge per iteration), AoS -> SoA
// conversion, and the calls into the C ABI.  No arithmetic of Solver::step() happens here.
#include "solver.hpp"

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace proj02 {

namespace {
void pose_to_wire(const NEPose& X, double* w) {
    w[0] = X.translation().x(); w[1] = X.translation().y();
    w[2] = X.linear()(0, 0); w[3] = X.linear()(1, 0);
}
NEPose wire_to_pose(const double* w) {
    NEPose X;
    X.translation() = la::Vec2f((float)w[0], (float)w[1]);
    X.linear()(0, 0) = (float)w[2]; X.linear()(0, 1) = (float)-w[3];
    X.linear()(1, 0) = (float)w[3]; X.linear()(1, 1) = (float)w[2];
    return X;
}
}  // namespace

Solver::Solver(const State& st, const BearingObservationVector& bear_obs, const OdometryObservationVector& odom_obs, const int& fixed_pose_id)
    : Solver(st, bear_obs, odom_obs, fixed_pose_id, SolverOptions()) {}

Solver::Solver(const State& st, const BearingObservationVector& bear_obs, const OdometryObservationVector& odom_obs, const int& fixed_pose_id,
               const SolverOptions& o)
    : state(st), bearing_observations(bear_obs), odometry_observations(odom_obs), fixed_pose_id_(fixed_pose_id) {
    N_ = 3 * state.number_of_poses() + 2 * state.number_of_landmarks();
    std::memset(&stats_, 0, sizeof(stats_));
    bos_default_options(&opt_);                 // kernel_threshold = 1, damping = 0.01f: the reference's defaults
    opt_.device = o.device;
    opt_.precision = o.fp32 ? BOS_PRECISION_F32 : BOS_PRECISION_F64;
    opt_.solver = o.solver;
    opt_.dense_max_dim = o.dense_max_dim;
    opt_.pcg_max_iters = o.pcg_max_iters;
    opt_.pcg_rtol = o.pcg_rtol;
    opt_.pcg_variant = o.pcg_variant;
    opt_.pcg_precond = o.pcg_precond;
    const int rc = bos_create(&opt_, &ctx_);
    if (rc != BOS_OK) throw std::runtime_error("proj02::Solver: bos_create failed with status " + std::to_string(rc) +
                                               " (no CUDA device? there is no CPU fallback)");
    try {
        upload_problem();
        upload_state();
    } catch (...) {
        bos_destroy(ctx_);
        ctx_ = nullptr;
        throw;
    }
}

Solver::~Solver() {
    if (ctx_) bos_destroy(ctx_);
}

void Solver::check(int rc, const char* what) const {
    if (rc != BOS_OK) throw std::runtime_error(std::string("proj02::Solver: ") + what + " failed (" + std::to_string(rc) + "): " + bos_last_error(ctx_));
}

void Solver::upload_problem() {
    const size_t Eb = bearing_observations.size(), Eo = odometry_observations.size();
    std::vector<int32_t> bp(Eb), bl(Eb), os(Eo), od(Eo);
    std::vector<double> bz(Eb), bom(Eb), oz(3 * Eo), oom(9 * Eo);
    for (size_t e = 0; e < Eb; e++) {
        const BearingObservation& b = bearing_observations[e];
        bp[e] = state.pose_stix(b.get_pose_id());       // std::out_of_range for an unknown id, as std::map::at in the reference
        bl[e] = state.landmark_stix(b.get_lm_id());
        bz[e] = b.get_bearing().angle();
        bom[e] = b.get_omega();
    }
    for (size_t e = 0; e < Eo; e++) {
        const OdometryObservation& o = odometry_observations[e];
        os[e] = state.pose_stix(o.get_source_id());
        od[e] = state.pose_stix(o.get_dest_id());
        const EPose z = o.get_transformation();
        const la::Mat3f om = o.get_omega();
        for (int k = 0; k < 3; k++) oz[3 * e + k] = z(k);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) oom[9 * e + 3 * i + j] = om(i, j);
    }
    const int fixed_stix = state.pose_stix(fixed_pose_id_);
    check(bos_upload_problem(ctx_, state.number_of_poses(), state.number_of_landmarks(), fixed_stix, (int64_t)Eb, bp.data(), bl.data(), bz.data(),
                             bom.data(), (int64_t)Eo, os.data(), od.data(), oz.data(), oom.data()),
          "bos_upload_problem");
}

void Solver::upload_state() {
    const int NP = state.number_of_poses(), NL = state.number_of_landmarks();
    if (3 * NP + 2 * NL != N_) throw std::runtime_error("proj02::Solver: the state changed size after construction");
    std::vector<double> P(4 * (size_t)NP), L(2 * (size_t)(NL > 0 ? NL : 1));
    for (int i = 0; i < NP; i++) pose_to_wire(state.pose_at(i), &P[4 * (size_t)i]);
    for (int j = 0; j < NL; j++) { L[2 * (size_t)j] = state.landmark_at(j).x(); L[2 * (size_t)j + 1] = state.landmark_at(j).y(); }
    check(bos_set_state(ctx_, P.data(), NL > 0 ? L.data() : nullptr), "bos_set_state");
    synced_version_ = state.version();
    device_state_valid_ = true;
}

void Solver::download_state() {
    const int NP = state.number_of_poses(), NL = state.number_of_landmarks();
    std::vector<double> P(4 * (size_t)NP), L(2 * (size_t)(NL > 0 ? NL : 1));
    check(bos_get_state(ctx_, P.data(), NL > 0 ? L.data() : nullptr), "bos_get_state");
    for (int i = 0; i < NP; i++) state.set_pose_at(i, wire_to_pose(&P[4 * (size_t)i]));
    for (int j = 0; j < NL; j++) state.set_landmark_at(j, LMPos((float)L[2 * (size_t)j], (float)L[2 * (size_t)j + 1]));
    synced_version_ = state.version();
}

void Solver::set_kernel_threshold(float kt) { check(bos_set_kernel_threshold(ctx_, kt), "bos_set_kernel_threshold"); }
void Solver::set_irls(bool on) { check(bos_set_robust_mode(ctx_, on ? BOS_ROBUST_IRLS : BOS_ROBUST_REFERENCE), "bos_set_robust_mode"); }
void Solver::set_damping_factor(float df) { check(bos_set_damping_factor(ctx_, df), "bos_set_damping_factor"); }

void Solver::step() { step(1, true); }

void Solver::step(int iterations, bool mirror_every_step) {
    if (!device_state_valid_ || state.version() != synced_version_) upload_state();
    for (int it = 0; it < iterations; it++) {
        check(bos_step(ctx_, &stats_), "bos_step");
        if (stats_.solver_status == 1)   // slam/solver.cpp:82-84: log and carry on
            std::cout << "Factorization failed: the reduced system is not positive definite" << std::endl;
        else if (stats_.solver_status == 2)
            std::cout << "PCG stopped at its iteration cap before reaching the tolerance (SolverOptions::pcg_max_iters)" << std::endl;
        if (mirror_every_step || it + 1 == iterations) download_state();
    }
}

bool Solver::step_lm() {
    if (!device_state_valid_ || state.version() != synced_version_) upload_state();
    int accepted = 0;
    check(bos_step_lm(ctx_, &stats_, nullptr, &accepted, nullptr), "bos_step_lm");
    download_state();
    return accepted != 0;
}

// ---- per-edge API ---------------------------------------------------------------------------------------------------------
void Solver::error_and_jacobian(const State& st, const BearingObservation& obs, float& error, SparseMatrixXf& jacobian) {
    const int N = 3 * st.number_of_poses() + 2 * st.number_of_landmarks();
    double P[4], L[2], z = obs.get_bearing().angle(), e = 0, J[5];
    pose_to_wire(st.get_pose_by_id(obs.get_pose_id()), P);
    const LMPos lm = st.get_landmark_by_id(obs.get_lm_id());
    L[0] = lm.x(); L[1] = lm.y();
    check(bos_eval_bearing_edges(&opt_, 1, P, L, &z, &e, J), "bos_eval_bearing_edges");
    error = (float)e;
    // column layout of the reference (slam/solver_jacobians.cpp:70-71): pose block at 3*pstix, landmark block at 3*NP + 2*lstix
    const int pc = 3 * st.pose_stix(obs.get_pose_id()), lc = 3 * st.number_of_poses() + 2 * st.landmark_stix(obs.get_lm_id());
    jacobian.resize(1, N);
    for (int k = 0; k < 3; k++) jacobian.coeffRef(0, pc + k) = (float)J[k];
    for (int k = 0; k < 2; k++) jacobian.coeffRef(0, lc + k) = (float)J[3 + k];
}

void Solver::error_and_jacobian(const State& st, const OdometryObservation& obs, EPose& error, SparseMatrixXf& jacobian) {
    const int N = 3 * st.number_of_poses() + 2 * st.number_of_landmarks();
    double S[4], D[4], z[3], e[3], J[18];
    pose_to_wire(st.get_pose_by_id(obs.get_source_id()), S);
    pose_to_wire(st.get_pose_by_id(obs.get_dest_id()), D);
    const EPose zt = obs.get_transformation();
    for (int k = 0; k < 3; k++) z[k] = zt(k);
    check(bos_eval_odometry_edges(&opt_, 1, S, D, z, e, J), "bos_eval_odometry_edges");
    error = EPose((float)e[0], (float)e[1], (float)e[2]);
    const int sc = 3 * st.pose_stix(obs.get_source_id()), dc = 3 * st.pose_stix(obs.get_dest_id());
    jacobian.resize(3, N);
    for (int i = 0; i < 3; i++)           // 18 entries including the explicit zeros (slam/solver_jacobians.cpp:139-165)
        for (int k = 0; k < 3; k++) {
            jacobian.coeffRef(i, sc + k) = (float)J[6 * i + k];
            jacobian.coeffRef(i, dc + k) = (float)J[6 * i + 3 + k];
        }
}

float Solver::normalized_angle(float angle) {   // slam/solver_jacobians.cpp:325-333, CV_PI / CV_2PI are doubles
    const double pi = 3.1415926535897932384626433832795, two_pi = 6.283185307179586476925286766559;
    while (angle < -pi) angle += two_pi;
    while (angle >= pi) angle -= two_pi;
    return angle;
}

float Solver::predict_bearing(const NEPose& pose, const LMPos& lm) {   // slam/solver_jacobians.cpp:301-305
    const LMPos g = pose.inverse() * lm;
    return std::atan2(g.y(), g.x());
}

EPose Solver::predict_odometry(const NEPose& src, const NEPose& dst) {   // slam/solver_jacobians.cpp:307-323
    const EPose s = t2v(src), d = t2v(dst);
    const la::Vec2f dt = src.rotation().transpose() * (dst.translation() - src.translation());
    return EPose(dt.x(), dt.y(), normalized_angle(d.z() - s.z()));
}

void Solver::error_and_numerical_jacobian(const State& st, const BearingObservation& obs, float& error, SparseMatrixXf& jacobian) {
    const NEPose pose = st.get_pose_by_id(obs.get_pose_id());
    const LMPos lm = st.get_landmark_by_id(obs.get_lm_id());
    const float zb = obs.get_bearing().smallestAngle(), eps = 0.001f;
    error = normalized_angle(predict_bearing(pose, lm) - zb);
    auto err_at = [&](const EPose& dp, const LMPos& dl) { return normalized_angle(predict_bearing(boxplus(pose, dp), lm + dl) - zb); };
    const int N = 3 * st.number_of_poses() + 2 * st.number_of_landmarks();
    const int pc = 3 * st.pose_stix(obs.get_pose_id()), lc = 3 * st.number_of_poses() + 2 * st.landmark_stix(obs.get_lm_id());
    jacobian.resize(1, N);
    for (int k = 0; k < 5; k++) {
        EPose dp(0, 0, 0); LMPos dl(0, 0);
        if (k < 3) dp(k) = eps; else dl(k - 3) = eps;
        const float d = (err_at(dp, dl) - err_at(dp * -1.f, dl * -1.f)) / (2 * eps);
        jacobian.coeffRef(0, k < 3 ? pc + k : lc + k - 3) = d;
    }
}

void Solver::error_and_numerical_jacobian(const State& st, const OdometryObservation& obs, EPose& error, SparseMatrixXf& jacobian) {
    const NEPose src = st.get_pose_by_id(obs.get_source_id()), dst = st.get_pose_by_id(obs.get_dest_id());
    const EPose z = obs.get_transformation();
    const float eps = 0.001f;
    auto err_at = [&](const EPose& ds, const EPose& dd) {
        EPose e = predict_odometry(boxplus(src, ds), boxplus(dst, dd)) - z;
        e.z() = normalized_angle(e.z());
        return e;
    };
    error = err_at(EPose(0, 0, 0), EPose(0, 0, 0));
    const int N = 3 * st.number_of_poses() + 2 * st.number_of_landmarks();
    const int sc = 3 * st.pose_stix(obs.get_source_id()), dc = 3 * st.pose_stix(obs.get_dest_id());
    jacobian.resize(3, N);
    for (int k = 0; k < 6; k++) {
        EPose ds(0, 0, 0), dd(0, 0, 0);
        if (k < 3) ds(k) = eps; else dd(k - 3) = eps;
        const EPose d = (err_at(ds, dd) - err_at(ds * -1.f, dd * -1.f)) * (1.f / (2 * eps));
        for (int i = 0; i < 3; i++) jacobian.coeffRef(i, k < 3 ? sc + k : dc + k - 3) = d(i);
    }
}

}  // namespace proj02

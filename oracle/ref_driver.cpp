// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C ABI around the REFERENCE'S OWN classes, compiled from the reference's own source files where they lie under /root/reference
// (oracle/Makefile, target `_ref`; nothing of the reference is copied into this repository):
//     framework/state.cpp  framework/observation.cpp  slam/solver.cpp  slam/solver_jacobians.cpp  slam/triangulation.cpp  utils/g2o_utils.cpp
// against oracle/eigen_standin (Eigen3 and OpenCV are not in this image).  The calls below are what the reference's main loop makes
// (executables/bearing_only_slam.cpp: parse_g2o, triangulate_landmarks, Solver ctor, step()); the getters read the solver's members.
// Used by tests/test_ref_build.py and tests/golden/make_ref_golden.py to pin bos_oracle.hpp; never by the product.
#include <fcntl.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Eigen/Sparse>

// the solver keeps H, b, H_nofixed, b_nofixed and the state's id tables private; this translation unit reads them
#define private public
#include "framework/state.hpp"
#include "framework/observation.hpp"
#include "slam/solver.hpp"
#undef private
#include "slam/triangulation.hpp"
#include "utils/g2o_utils.hpp"
#include "utils/draw_utils.hpp"

namespace proj02 {
// framework/state.cpp:98-108 (State::draw) references two drawing functions of utils/draw_utils.cpp (OpenCV, out of scope): link stubs
void draw_poses(RGBImage&, const NEPoseVector&, const float&) {}
void draw_landmarks(RGBImage&, const LMPosVector&, const float&) {}
}  // namespace proj02

using namespace proj02;

namespace {
struct Ref {
    State state;
    BearingObservationVector bearings;
    OdometryObservationVector odoms;
    int fixed_pose_id = -1;
    float bound = 0;
    Solver* solver = nullptr;
    ~Ref() { delete solver; }
    State& st() { return solver ? solver->state : state; }
};
Ref* R(void* h) { return static_cast<Ref*>(h); }

struct Quiet {   // the reference prints its warnings to std::cout (flushed by std::endl): point fd 1 at /dev/null meanwhile
    int saved;
    Quiet() {
        std::cout.flush();
        fflush(stdout);
        saved = dup(1);
        const int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) { dup2(nul, 1); close(nul); }
    }
    ~Quiet() {
        std::cout.flush();
        fflush(stdout);
        if (saved >= 0) { dup2(saved, 1); close(saved); }
    }
};
}  // namespace

extern "C" {

void* ref_new() { return new Ref(); }
void ref_free(void* h) { delete R(h); }

int ref_load_g2o(void* h, const char* path) {
    std::ifstream probe(path);
    if (!probe.good()) return 1;
    Quiet q;
    parse_g2o(std::string(path), R(h)->state, R(h)->bearings, R(h)->odoms, R(h)->fixed_pose_id, R(h)->bound);
    return 0;
}
// problems given as arrays go through the same public calls the parser makes (utils/g2o_utils.cpp:48, 68, 112, 124)
void ref_add_poses(void* h, int n, const int* ids, const float* xyt) {
    for (int i = 0; i < n; i++) R(h)->state.add_pose(xyt[3 * i], xyt[3 * i + 1], xyt[3 * i + 2], ids[i]);
}
void ref_add_landmarks(void* h, int n, const int* ids, const float* xy) {
    for (int i = 0; i < n; i++) R(h)->state.add_landmark(xy[2 * i], xy[2 * i + 1], ids[i]);
}
void ref_add_bearings(void* h, int n, const int* pose_id, const int* lm_id, const float* z, const float* omega) {
    for (int i = 0; i < n; i++) {
        if (omega) R(h)->bearings.emplace_back(pose_id[i], lm_id[i], z[i], omega[i]);
        else R(h)->bearings.emplace_back(pose_id[i], lm_id[i], z[i]);
    }
}
void ref_add_odometry(void* h, int n, const int* src, const int* dst, const float* z, const float* omega9) {
    for (int i = 0; i < n; i++) {
        Eigen::Matrix3f om;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) om(r, c) = omega9[9 * i + 3 * r + c];
        R(h)->odoms.emplace_back(src[i], dst[i], z[3 * i], z[3 * i + 1], z[3 * i + 2], om);
    }
}
void ref_set_fixed(void* h, int id) { R(h)->fixed_pose_id = id; }

int ref_triangulate(void* h) {
    Quiet q;
    try {
        triangulate_landmarks(R(h)->state, R(h)->bearings);
    } catch (const std::out_of_range&) {
        return 1;
    }
    return 0;
}

// out[0..5] = NP, NL, Eb, Eo, fixed pose id, N
void ref_counts(void* h, int* out) {
    State& s = R(h)->st();
    out[0] = s.number_of_poses();
    out[1] = s.number_of_landmarks();
    out[2] = (int)R(h)->bearings.size();
    out[3] = (int)R(h)->odoms.size();
    out[4] = R(h)->fixed_pose_id;
    out[5] = 3 * out[0] + 2 * out[1];
}
float ref_bound(void* h) { return R(h)->bound; }
void ref_get_ids(void* h, int* pose_ids, int* lm_ids) {
    State& s = R(h)->st();
    for (size_t i = 0; i < s.pose_stix_to_id.size(); i++) pose_ids[i] = s.pose_stix_to_id[i];
    for (size_t i = 0; i < s.lm_stix_to_id.size(); i++) lm_ids[i] = s.lm_stix_to_id[i];
}
// poses as (tx, ty, R00 = cos, R10 = sin), landmarks as (x, y): the raw members, no t2v in between
void ref_get_state(void* h, float* poses_xycs, float* lms_xy) {
    State& s = R(h)->st();
    for (size_t i = 0; i < s.poses.size(); i++) {
        poses_xycs[4 * i + 0] = s.poses[i].translation()(0);
        poses_xycs[4 * i + 1] = s.poses[i].translation()(1);
        poses_xycs[4 * i + 2] = s.poses[i].linear()(0, 0);
        poses_xycs[4 * i + 3] = s.poses[i].linear()(1, 0);
    }
    for (size_t j = 0; j < s.landmarks.size(); j++) {
        lms_xy[2 * j] = s.landmarks[j](0);
        lms_xy[2 * j + 1] = s.landmarks[j](1);
    }
}
void ref_get_state_xyt(void* h, float* poses_xyt) {
    State& s = R(h)->st();
    for (size_t i = 0; i < s.poses.size(); i++) {
        EPose e = t2v(s.poses[i]);
        for (int k = 0; k < 3; k++) poses_xyt[3 * i + k] = e(k);
    }
}

int ref_solver_init(void* h, int fixed_id) {
    delete R(h)->solver;
    R(h)->solver = nullptr;
    if (fixed_id < 0) fixed_id = R(h)->fixed_pose_id >= 0 ? R(h)->fixed_pose_id : R(h)->state.default_pose_id();   // bearing_only_slam.cpp
    try {
        R(h)->solver = new Solver(R(h)->state, R(h)->bearings, R(h)->odoms, fixed_id);
    } catch (const std::out_of_range&) {
        return 1;
    }
    R(h)->fixed_pose_id = fixed_id;
    return 0;
}
void ref_set_params(void* h, float kernel_threshold, float damping) {
    R(h)->solver->set_kernel_threshold(kernel_threshold);
    R(h)->solver->set_damping_factor(damping);
}
int ref_step(void* h) {
    Quiet q;
    try {
        R(h)->solver->step();
    } catch (const std::out_of_range&) {
        return 1;
    }
    return R(h)->solver->sparse_system_solver.info() == Eigen::Success ? 0 : 2;   // 2 = the reference's "not SPD" warning was printed
}

// ---- what the last step() left in the solver's members ----
static const SparseMatrixXf& which_H(void* h, int nofixed) { return nofixed ? R(h)->solver->H_nofixed : R(h)->solver->H; }
long ref_H_nnz(void* h, int nofixed) { return which_H(h, nofixed).nonZeros(); }
void ref_get_H(void* h, int nofixed, int* colptr, int* rowidx, float* val) {
    const SparseMatrixXf& A = which_H(h, nofixed);
    int p = 0;
    for (int j = 0; j < A.outerSize(); j++) {
        colptr[j] = p;
        for (SparseMatrixXf::InnerIterator it(A, j); it; ++it, ++p) {
            rowidx[p] = it.row();
            val[p] = it.value();
        }
    }
    colptr[A.outerSize()] = p;
}
void ref_get_b(void* h, int nofixed, float* b) {
    const Eigen::VectorXf& v = nofixed ? R(h)->solver->b_nofixed : R(h)->solver->b;
    for (int i = 0; i < v.size(); i++) b[i] = v(i);
}

// ---- per-edge terms through the public error_and_jacobian (analytic = 0, numeric = 1), at the solver's current state ----
// eb[Eb], jb[Eb][5] = (pose 1x3 | landmark 1x2); eo[Eo][3], jo[Eo][18] = 3 x 6 row-major (source 3x3 | destination 3x3)
void ref_edge_terms(void* h, int numeric, float* eb, float* jb, float* eo, float* jo) {
    Solver& S = *R(h)->solver;
    const int NPp = S.state.number_of_poses();
    for (size_t e = 0; e < S.bearing_observations.size(); e++) {
        const BearingObservation& obs = S.bearing_observations[e];
        float err;
        SparseMatrixXf J;
        J.resize(1, S.N);
        if (numeric) S.error_and_numerical_jacobian(S.state, obs, err, J);
        else S.error_and_jacobian(S.state, obs, err, J);
        const int pc = 3 * S.state.pose_stix(obs.get_pose_id()), lc = 3 * NPp + 2 * S.state.landmark_stix(obs.get_lm_id());
        eb[e] = err;
        for (int k = 0; k < 3; k++) jb[5 * e + k] = J.coeff(0, pc + k);
        for (int k = 0; k < 2; k++) jb[5 * e + 3 + k] = J.coeff(0, lc + k);
    }
    for (size_t e = 0; e < S.odometry_observations.size(); e++) {
        const OdometryObservation& obs = S.odometry_observations[e];
        EPose err;
        SparseMatrixXf J;
        J.resize(3, S.N);
        if (numeric) S.error_and_numerical_jacobian(S.state, obs, err, J);
        else S.error_and_jacobian(S.state, obs, err, J);
        const int sc = 3 * S.state.pose_stix(obs.get_source_id()), dc = 3 * S.state.pose_stix(obs.get_dest_id());
        for (int r = 0; r < 3; r++) {
            eo[3 * e + r] = err(r);
            for (int k = 0; k < 3; k++) {
                jo[18 * e + 6 * r + k] = J.coeff(r, sc + k);
                jo[18 * e + 6 * r + 3 + k] = J.coeff(r, dc + k);
            }
        }
    }
}

// ---- the small public functions the reference's tests print (tests/solver_stuff.cpp:25-38, 93-114) ----
float ref_predict_bearing(void* h, float x, float y, float th, float lx, float ly) {
    return R(h)->solver->predict_bearing(v2t(EPose(x, y, th)), LMPos(lx, ly));
}
void ref_predict_odometry(void* h, const float* s, const float* d, float* out) {
    EPose p = R(h)->solver->predict_odometry(v2t(EPose(s[0], s[1], s[2])), v2t(EPose(d[0], d[1], d[2])));
    for (int k = 0; k < 3; k++) out[k] = p(k);
}
float ref_normalized_angle(void* h, float a) { return R(h)->solver->normalized_angle(a); }
float ref_smallest_angle(float a) { return Rotation2f(a).smallestAngle(); }
// boxplus(v2t(x), d) -> t2v (framework/state.hpp:11-13)
void ref_boxplus(const float* xyt, const float* d, float* out) {
    EPose e = t2v(boxplus(v2t(EPose(xyt[0], xyt[1], xyt[2])), EPose(d[0], d[1], d[2])));
    for (int k = 0; k < 3; k++) out[k] = e(k);
}

}  // extern "C"

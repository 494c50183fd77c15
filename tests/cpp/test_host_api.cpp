// test_host_api.cpp -- exercises the C++ host mirror of the reference's API the way the reference's own manual tests do
// (tests/state_test.cpp, tests/observation_test.cpp, tests/solver_stuff.cpp, tests/triangulation_test.cpp,
// executables/bearing_only_slam.cpp), but with assertions / machine-readable output instead of OpenCV windows.
//   test_host_api cpu  <initial_guess.g2o>                       host-only checks (no device)
//   test_host_api gpu  <initial_guess.g2o> <ground_truth.g2o> <iters>   full flow on the device, prints key=value lines
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "slam/solver.hpp"
#include "slam/triangulation.hpp"
#include "utils/g2o_utils.hpp"

using namespace proj02;

static int failures = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { std::printf("CHECK FAILED %s:%d %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

static const float PI = 3.14159265358979323846f;

static void cpu_checks(const char* ig) {
    // tests/state_test.cpp:10-18: non-contiguous ids; a duplicate landmark id overwrites the map entry, the vector keeps both
    State state(5, 5);
    state.add_pose(100, 300, PI / 4, 897);
    state.add_pose(250, 378, PI, 357);
    state.add_pose(400, 128, -PI / 3, 205);
    state.add_landmark(200, 200, 35);
    state.add_landmark(400, 400, 35);
    CHECK(state.number_of_poses() == 3 && state.number_of_landmarks() == 2);
    CHECK(state.pose_stix(897) == 0 && state.pose_stix(357) == 1 && state.pose_stix(205) == 2);
    CHECK(state.landmark_stix(35) == 1 && state.get_landmark_by_id(35).x() == 400.f);
    CHECK(state.default_pose_id() == 897);
    bool threw = false;
    try { state.get_pose_by_id(1); } catch (const std::out_of_range&) { threw = true; }
    CHECK(threw);
    // t2v / v2t / boxplus (framework/definitions.hpp:39-53, state.hpp:11-13)
    const EPose p = t2v(state.get_pose_by_id(205));
    CHECK(std::fabs(p.x() - 400.f) < 1e-4f && std::fabs(p.z() + PI / 3) < 1e-6f);
    const NEPose moved = boxplus(v2t(EPose(1, 2, 0.5f)), EPose(0.1f, -0.2f, 0.25f));   // v2t(delta) * X
    const EPose mv = t2v(moved);
    const float c = std::cos(0.25f), s = std::sin(0.25f);
    CHECK(std::fabs(mv.x() - (c * 1 - s * 2 + 0.1f)) < 1e-6f && std::fabs(mv.y() - (s * 1 + c * 2 - 0.2f)) < 1e-6f && std::fabs(mv.z() - 0.75f) < 1e-6f);
    la::VectorXf dx(3 * 3 + 2 * 2);
    dx(2) = 0.1f; dx(9) = 1.f; dx(12) = -2.f;
    state.apply_boxplus(dx);
    CHECK(std::fabs(t2v(state.get_pose_by_id(897)).z() - (PI / 4 + 0.1f)) < 1e-6f);
    CHECK(state.landmark_at(0).x() == 201.f && state.landmark_at(1).y() == 398.f);
    // tests/observation_test.cpp: bearing = world angle - pose theta, stored un-normalised
    BearingObservation bo(897, 35, 7.0f);
    CHECK(bo.get_bearing().angle() == 7.0f && std::fabs(bo.get_bearing().smallestAngle() - (7.0f - 2 * PI)) < 1e-6f && bo.get_omega() == 1.f);
    la::Mat3f om = la::Mat3f::Identity();
    om(0, 1) = om(1, 0) = 2.f;
    OdometryObservation oo(1, 2, 0.5f, 0.25f, -0.1f, om);
    CHECK(oo.get_source_id() == 1 && oo.get_dest_id() == 2 && oo.get_transformation().y() == 0.25f && oo.get_omega()(1, 0) == 2.f);
    CHECK(oo.get_omega_sparse().nonZeros() == 9);
    // loader (utils/g2o_utils.cpp:10-146)
    State st(300, 200);
    BearingObservationVector be;
    OdometryObservationVector od;
    int fixed = 0;
    float bound = 0;
    parse_g2o(ig, st, be, od, fixed, bound);
    std::printf("poses=%d landmarks=%d bearings=%zu odometries=%zu fixed=%d bound=%.9g\n", st.number_of_poses(), st.number_of_landmarks(), be.size(),
                od.size(), fixed, bound);
    if (st.number_of_poses() > 0) {
        const EPose p0 = t2v(st.pose_at(0));
        std::printf("pose0=%d %.9g %.9g %.9g\n", st.pose_id_at(0), p0.x(), p0.y(), p0.z());
        std::printf("edge0=%d %d %.9g\n", be[0].get_pose_id(), be[0].get_lm_id(), be[0].get_bearing().angle());
        if (!od.empty()) std::printf("odom0=%d %d %.9g %.9g %.9g %.9g %.9g\n", od[0].get_source_id(), od[0].get_dest_id(), od[0].get_transformation().x(),
                                     od[0].get_transformation().y(), od[0].get_transformation().z(), od[0].get_omega()(0, 0), od[0].get_omega()(2, 2));
        // write -> parse round trip of the extension writer
        const std::string tmp = std::string(ig) + ".roundtrip";
        CHECK(write_g2o(tmp, st, be, od, fixed));
        State st2;
        BearingObservationVector be2;
        OdometryObservationVector od2;
        int fixed2;
        float bound2;
        parse_g2o(tmp, st2, be2, od2, fixed2, bound2);
        CHECK(st2.number_of_poses() == st.number_of_poses() && be2.size() == be.size() && od2.size() == od.size() && fixed2 == fixed);
        bool same = true;
        for (int i = 0; i < st.number_of_poses(); i++) same = same && t2v(st2.pose_at(i)).x() == t2v(st.pose_at(i)).x() && st2.pose_id_at(i) == st.pose_id_at(i);
        for (size_t e = 0; e < be.size(); e++) same = same && be2[e].get_bearing().angle() == be[e].get_bearing().angle();
        CHECK(same);
        std::remove(tmp.c_str());
    }
    // a missing file: two warnings, fixed = -1, bound = 3 (utils/g2o_utils.cpp:11-12,135-143)
    State empty;
    BearingObservationVector be3;
    int fixed3 = 7;
    float bound3 = 9;
    parse_g2o("/nonexistent/file.g2o", empty, be3, fixed3, bound3);
    CHECK(empty.number_of_poses() == 0 && fixed3 == -1 && bound3 == 3.f);
}

static void gpu_checks(const char* ig, const char* gt, int iters) {
    State state(300, 200);
    BearingObservationVector be;
    OdometryObservationVector od;
    int fixed;
    float bound;
    parse_g2o(ig, state, be, od, fixed, bound);
    if (fixed < 0) fixed = state.default_pose_id();
    triangulate_landmarks(state, be);
    std::printf("landmarks_after_triangulation=%d first_lm_id=%d last_lm_id=%d\n", state.number_of_landmarks(), state.landmark_id_at(0),
                state.landmark_id_at(state.number_of_landmarks() - 1));
    for (int j = 0; j < state.number_of_landmarks(); j++)
        std::printf("tri %d %.9g %.9g\n", state.landmark_id_at(j), state.landmark_at(j).x(), state.landmark_at(j).y());
    {
        Solver solver(state, be, od, fixed);
        // tests/solver_stuff.cpp:25-38 known answers
        const float a[7] = {solver.predict_bearing(v2t(EPose(0, 0, 0)), LMPos(1, 0)),  solver.predict_bearing(v2t(EPose(0, 0, 0)), LMPos(0, 1)),
                            solver.predict_bearing(v2t(EPose(0, 0, 0)), LMPos(-1, 0)), solver.predict_bearing(v2t(EPose(0, 0, 0)), LMPos(0, -1)),
                            solver.predict_bearing(v2t(EPose(0, 0, 0)), LMPos(1, 1)),  solver.predict_bearing(v2t(EPose(0, 0, PI / 2)), LMPos(1, 1)),
                            solver.predict_bearing(v2t(EPose(0, 0, PI)), LMPos(1, 0))};
        CHECK(a[0] == 0.f && std::fabs(a[1] - PI / 2) < 1e-6f && std::fabs(std::fabs(a[2]) - PI) < 1e-6f && std::fabs(a[3] + PI / 2) < 1e-6f);
        CHECK(std::fabs(a[4] - PI / 4) < 1e-6f && std::fabs(a[5] + PI / 4) < 1e-6f && std::fabs(std::fabs(a[6]) - PI) < 1e-6f);
        CHECK(solver.normalized_angle(3.5f) < 0 && solver.normalized_angle(-3.5f) > 0 && solver.normalized_angle(1.f) == 1.f);
        // tests/solver_stuff.cpp:93-114: predict_odometry ~ measurement on the dead-reckoned initial guess
        double worst = 0;
        const int idx[8] = {0, 10, 42, 111, 128, 163, 222, 255};
        for (int k = 0; k < 8; k++)
            if (idx[k] < (int)od.size()) {
                const EPose pr = solver.predict_odometry(state.get_pose_by_id(od[idx[k]].get_source_id()), state.get_pose_by_id(od[idx[k]].get_dest_id()));
                const EPose z = od[idx[k]].get_transformation();
                worst = std::fmax(worst, std::fmax(std::fabs(pr.x() - z.x()), std::fmax(std::fabs(pr.y() - z.y()), std::fabs(solver.normalized_angle(pr.z() - z.z())))));
            }
        std::printf("predict_odometry_worst=%.6g\n", worst);
        // tests/solver_stuff.cpp:117-163: analytic (device) vs numeric (host) odometry Jacobian on the triangulated initial guess
        float hs = 0, hm = 0, ts = 0, tm = 0;
        const size_t nod = od.size() < 60 ? od.size() : 60;
        for (size_t e = 0; e < nod; e++) {
            EPose err_a, err_n;
            SparseMatrixXf ja, jn;
            solver.error_and_jacobian(state, od[e], err_a, ja);
            solver.error_and_numerical_jacobian(state, od[e], err_n, jn);
            CHECK(ja.nonZeros() == 18 && ja.rows() == 3 && ja.cols() == 3 * state.number_of_poses() + 2 * state.number_of_landmarks());
            CHECK(std::fabs(err_a.x() - err_n.x()) < 1e-5f && std::fabs(solver.normalized_angle(err_a.z() - err_n.z())) < 1e-5f);
            const SparseMatrixXf d = (ja - jn).cwiseAbs();
            hs = std::fmax(hs, d.sum()); hm = std::fmax(hm, d.coeffs().maxCoeff());
            ts += d.sum(); tm += d.coeffs().maxCoeff();
        }
        if (nod) std::printf("odom_jacobian highest_sum=%.6g highest_max=%.6g average_sum=%.6g average_max=%.6g n=%zu\n", hs, hm, ts / nod, tm / nod, nod);
        // the reference's call sequence: step() x K, state read back after every step
        for (int it = 0; it < iters; it++) {
            solver.step();
            const bos_stats& s = solver.last_stats();
            std::printf("it %d chi2_bearing=%.12e chi2_odometry=%.12e over_bearing=%lld delta_inf=%.9e status=%d\n", it, s.chi2_bearing, s.chi2_odometry,
                        (long long)s.over_bearing, s.delta_inf, s.solver_status);
        }
        for (int i = 0; i < solver.state.number_of_poses(); i++) {
            const EPose p = t2v(solver.state.pose_at(i));
            std::printf("final_pose %d %.9g %.9g %.9g\n", solver.state.pose_id_at(i), p.x(), p.y(), p.z());
        }
        for (int j = 0; j < solver.state.number_of_landmarks(); j++)
            std::printf("final_lm %d %.9g %.9g\n", solver.state.landmark_id_at(j), solver.state.landmark_at(j).x(), solver.state.landmark_at(j).y());
        // a caller-side modification of solver.state is picked up by the next step (version counter)
        const EPose before = t2v(solver.state.pose_at(solver.state.number_of_poses() - 1));
        la::VectorXf nudge(3 * (size_t)solver.state.number_of_poses() + 2 * (size_t)solver.state.number_of_landmarks());
        nudge(3 * (size_t)(solver.state.number_of_poses() - 1)) = 0.5f;
        solver.state.apply_boxplus(nudge);
        solver.step();
        const EPose after = t2v(solver.state.pose_at(solver.state.number_of_poses() - 1));
        std::printf("nudge_recovery=%.6g chi2_after_nudge=%.9e\n", std::fabs(after.x() - before.x()), solver.last_stats().chi2_bearing + solver.last_stats().chi2_odometry);
        // unknown ids throw like std::map::at
        bool threw = false;
        try {
            BearingObservationVector bad = be;
            bad.emplace_back(123456789, be[0].get_lm_id(), 0.f);
            Solver s2(state, bad, od, fixed);
        } catch (const std::out_of_range&) { threw = true; }
        CHECK(threw);
    }
    // tests/solver_stuff.cpp:42-89: analytic vs numeric bearing Jacobian on the ground-truth state
    State gts;
    BearingObservationVector gbe;
    OdometryObservationVector god;
    int gfixed;
    float gbound;
    parse_g2o(gt, gts, gbe, god, gfixed, gbound);
    if (gts.number_of_landmarks() > 0) {
        OdometryObservationVector none;
        Solver s(gts, gbe, none, gfixed < 0 ? gts.default_pose_id() : gfixed);
        float hs = 0, hm = 0, ts = 0, tm = 0;
        const size_t nb = gbe.size() < 300 ? gbe.size() : 300;
        for (size_t e = 0; e < nb; e++) {
            float ea, en;
            SparseMatrixXf ja, jn;
            s.error_and_jacobian(gts, gbe[e], ea, ja);
            s.error_and_numerical_jacobian(gts, gbe[e], en, jn);
            CHECK(ja.nonZeros() == 5 && std::fabs(ea - en) < 1e-5f);
            const SparseMatrixXf d = (ja - jn).cwiseAbs();
            hs = std::fmax(hs, d.sum()); hm = std::fmax(hm, d.coeffs().maxCoeff());
            ts += d.sum(); tm += d.coeffs().maxCoeff();
        }
        std::printf("bearing_jacobian highest_sum=%.6g highest_max=%.6g average_sum=%.6g average_max=%.6g n=%zu\n", hs, hm, ts / nb, tm / nb, nb);
    }
}

int main(int argc, char** argv) {
    if (argc >= 3 && !std::strcmp(argv[1], "cpu")) cpu_checks(argv[2]);
    else if (argc >= 5 && !std::strcmp(argv[1], "gpu")) {
        try { gpu_checks(argv[2], argv[3], std::atoi(argv[4])); } catch (const std::exception& e) { std::printf("EXCEPTION %s\n", e.what()); failures++; }
    } else { std::printf("usage: test_host_api cpu <ig.g2o> | gpu <ig.g2o> <gt.g2o> <iters>\n"); return 2; }
    std::printf("failures=%d\n", failures);
    return failures ? 1 : 0;
}

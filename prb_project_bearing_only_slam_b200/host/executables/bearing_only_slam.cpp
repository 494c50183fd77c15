// bearing_only_slam.cpp -- headless harness replacing the reference's interactive OpenCV loop
// (executables/bearing_only_slam.cpp:40-115).  Same call sequence: parse_g2o -> default fixed pose -> triangulate_landmarks
// -> Solver -> step() x K; instead of drawing it prints per-iteration chi2 / |dx| and can dump the final state as g2o.
//
//   bearing_only_slam <dataset.g2o> [--iters K=50] [--out final.g2o] [--fp32] [--solver auto|dense|pcg] [--device D] [--quiet]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "../slam/solver.hpp"
#include "../slam/triangulation.hpp"
#include "../utils/g2o_utils.hpp"

using namespace proj02;

int main(int argc, char** argv) {
    if (argc < 2) {
        std::cout << "usage: bearing_only_slam <dataset_fname> [--iters K] [--out final.g2o] [--fp32] [--solver auto|dense|pcg] [--device D] [--quiet]" << std::endl;
        return 1;
    }
    int iters = 50;   // one Tab press of the reference (executables/bearing_only_slam.cpp:95)
    std::string out;
    bool quiet = false;
    SolverOptions opt;
    for (int a = 2; a < argc; a++) {
        const std::string s = argv[a];
        if (s == "--iters" && a + 1 < argc) iters = std::atoi(argv[++a]);
        else if (s == "--out" && a + 1 < argc) out = argv[++a];
        else if (s == "--fp32") opt.fp32 = true;
        else if (s == "--quiet") quiet = true;
        else if (s == "--device" && a + 1 < argc) opt.device = std::atoi(argv[++a]);
        else if (s == "--solver" && a + 1 < argc) {
            const std::string v = argv[++a];
            opt.solver = v == "dense" ? BOS_SOLVER_DENSE_CHOLESKY : v == "pcg" ? BOS_SOLVER_PCG : BOS_SOLVER_AUTO;
        } else { std::cout << "unknown option " << s << std::endl; return 1; }
    }
    State state(300, 200);
    BearingObservationVector bearing_observations;
    bearing_observations.reserve(1800);
    OdometryObservationVector odometry_observations;
    odometry_observations.reserve(300);
    int fixed_pose_id;
    float bound = 0;
    parse_g2o(argv[1], state, bearing_observations, odometry_observations, fixed_pose_id, bound);
    if (state.number_of_poses() == 0) return 2;
    if (fixed_pose_id < 0) fixed_pose_id = state.default_pose_id();
    try {
        triangulate_landmarks(state, bearing_observations, opt.device, opt.fp32);
        Solver solver(state, bearing_observations, odometry_observations, fixed_pose_id, opt);
        std::printf("poses %d landmarks %d bearing_edges %zu odometry_edges %zu fixed_pose %d bound %g\n", state.number_of_poses(),
                    state.number_of_landmarks(), bearing_observations.size(), odometry_observations.size(), fixed_pose_id, bound);
        const auto t0 = std::chrono::steady_clock::now();
        for (int it = 0; it < iters; it++) {
            solver.step();
            const bos_stats& s = solver.last_stats();
            if (!quiet)
                std::printf("it %3d chi2_bearing %.9e chi2_odometry %.9e over %lld %lld dx_inf %.6e solver %s pcg_it %d ms %.3f\n", it, s.chi2_bearing,
                            s.chi2_odometry, (long long)s.over_bearing, (long long)s.over_odometry, s.delta_inf,
                            s.solver_used == BOS_SOLVER_PCG ? "pcg" : "dense", s.pcg_iterations, s.ms_linearize + s.ms_allreduce + s.ms_solve + s.ms_update);
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const bos_stats& s = solver.last_stats();
        std::printf("done iterations %d wall_s %.6f iterations_per_s %.3f final_chi2 %.9e\n", iters, sec, iters / (sec > 0 ? sec : 1), s.chi2_bearing + s.chi2_odometry);
        if (!out.empty() && !write_g2o(out, solver.state, bearing_observations, odometry_observations, fixed_pose_id)) {
            std::cout << "cannot write " << out << std::endl;
            return 3;
        }
    } catch (const std::exception& e) {
        std::cout << "error: " << e.what() << std::endl;
        return 4;
    }
    return 0;
}

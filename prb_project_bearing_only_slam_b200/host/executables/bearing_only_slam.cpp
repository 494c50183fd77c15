// bearing_only_slam.cpp -- headless harness replacing the reference's interactive OpenCV loop
// (executables/bearing_only_slam.cpp:40-115).  Same call sequence: parse_g2o -> default fixed pose -> triangulate_landmarks
// -> Solver -> step() x K; instead of drawing it prints per-iteration chi2 / |dx| and can dump the final state as g2o.
//
//   bearing_only_slam <dataset.g2o> [--iters K=50] [--out final.g2o] [--fp32] [--solver auto|dense|pcg] [--device D] [--quiet]
//   bearing_only_slam --synth NP NL EDGES [same options]      a synthetic world of that size instead of a file (BASELINE configs 3, 4)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../slam/solver.hpp"
#include "../slam/triangulation.hpp"
#include "../utils/g2o_utils.hpp"
#include "bos_synth.h"   // synth/: the input generator is its own library, not part of libbos_b200.so

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
    int synth_np = 0, synth_nl = 0;
    long long synth_e = 0;
    int first_opt = 2;
    if (std::string(argv[1]) == "--synth") {
        if (argc < 5) { std::cout << "--synth needs NP NL EDGES" << std::endl; return 1; }
        synth_np = std::atoi(argv[2]); synth_nl = std::atoi(argv[3]); synth_e = std::atoll(argv[4]);
        first_opt = 5;
    }
    for (int a = first_opt; a < argc; a++) {
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
    if (synth_np > 0) {
        // the synthetic world generator (synth/libbos_synth.so), fed through the same State / observation API a file would go through
        bos_synth_spec spec;
        bos_synth_default_spec(&spec);
        spec.n_poses = synth_np; spec.n_landmarks = synth_nl; spec.target_bearing_edges = synth_e;
        bos_synth* w = nullptr;
        if (bos_synth_create(&spec, &w) != 0) { std::cout << "cannot generate the synthetic world" << std::endl; return 2; }
        int64_t cnt[4];
        bos_synth_counts(w, cnt);
        std::vector<int32_t> pid(cnt[0]), bp(cnt[2]), bl(cnt[2]), os(cnt[3]), od(cnt[3]);
        std::vector<double> xyt(3 * cnt[0]), bz(cnt[2]), oz(3 * cnt[3]), oom(9 * cnt[3]);
        bos_synth_get(w, pid.data(), xyt.data(), nullptr, nullptr, nullptr, bp.data(), bl.data(), bz.data(), os.data(), od.data(), oz.data(), oom.data());
        bos_synth_destroy(w);
        for (int64_t i = 0; i < cnt[0]; i++) state.add_pose((float)xyt[3 * i], (float)xyt[3 * i + 1], (float)xyt[3 * i + 2], pid[i]);
        bearing_observations.reserve(cnt[2]);
        for (int64_t e = 0; e < cnt[2]; e++) bearing_observations.emplace_back(bp[e], bl[e], (float)bz[e]);
        odometry_observations.reserve(cnt[3]);
        for (int64_t e = 0; e < cnt[3]; e++) {
            la::Mat3f om;
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) om(a, b) = (float)oom[9 * e + 3 * a + b];
            odometry_observations.emplace_back(os[e], od[e], (float)oz[3 * e], (float)oz[3 * e + 1], (float)oz[3 * e + 2], om);
        }
        fixed_pose_id = -1;
    } else {
        parse_g2o(argv[1], state, bearing_observations, odometry_observations, fixed_pose_id, bound);
    }
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

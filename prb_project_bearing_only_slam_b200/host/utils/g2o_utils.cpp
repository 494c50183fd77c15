// g2o_utils.cpp -- same observable behaviour as utils/g2o_utils.cpp:10-146 of the reference: unknown line types print
// "Unrecognized <token>", a missing file yields the two warnings, malformed numbers throw from std::stoi / std::stof.
#include "g2o_utils.hpp"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

namespace proj02 {

void parse_g2o(std::string fname, State& state, BearingObservationVector& bearings, int& fixed_pose_id, float& bound) {
    OdometryObservationVector ignored;
    parse_g2o(fname, state, bearings, ignored, fixed_pose_id, bound);
}

namespace {
struct Tokens {
    std::istringstream in;
    explicit Tokens(const std::string& line) : in(line) {}
    std::string word() { std::string t; in >> t; return t; }
    int i() { return std::stoi(word()); }
    float f() { return std::stof(word()); }
};
}  // namespace

void parse_g2o(std::string fname, State& state, BearingObservationVector& bearings, OdometryObservationVector& odometries,
               int& fixed_pose_id, float& bound) {
    bound = 0;
    fixed_pose_id = -1;
    std::ifstream f(fname);
    std::string line;
    auto grow = [&](float v) { if (std::abs(v) > bound) bound = std::abs(v); };
    while (std::getline(f, line)) {
        Tokens tk(line);
        const std::string kind = tk.word();
        if (kind == "VERTEX_SE2") {
            const int id = tk.i();
            const float x = tk.f(), y = tk.f(), th = tk.f();
            grow(x); grow(y);
            state.add_pose(x, y, th, id);
        } else if (kind == "VERTEX_XY") {
            const int id = tk.i();
            const float x = tk.f(), y = tk.f();
            grow(x); grow(y);
            state.add_landmark(x, y, id);
        } else if (kind == "FIX") {
            fixed_pose_id = tk.i();
        } else if (kind == "EDGE_SE2") {
            const int src = tk.i(), dst = tk.i();
            const float x = tk.f(), y = tk.f(), th = tk.f();
            la::Mat3f om;
            om(0, 0) = tk.f();
            om(0, 1) = om(1, 0) = tk.f();
            om(0, 2) = om(2, 0) = tk.f();
            om(1, 1) = tk.f();
            om(1, 2) = om(2, 1) = tk.f();
            om(2, 2) = tk.f();
            odometries.emplace_back(src, dst, x, y, th, om);
        } else if (kind == "EDGE_BEARING_SE2_XY") {
            const int pid = tk.i(), lid = tk.i();
            const float bearing = tk.f();   // the information value that follows is ignored (data/README.txt:12)
            bearings.emplace_back(pid, lid, bearing);
        } else if (kind.empty()) {
            // blank line
        } else {
            std::cout << "Unrecognized " << kind << std::endl;
        }
    }
    bound += 3;
    if (state.number_of_poses() == 0) std::cout << "Warning: no poses found. Stuff is likely to break." << std::endl;
    if (bearings.size() == 0) std::cout << "Warning: no bearing observations found. Stuff is likely to break." << std::endl;
}

bool write_g2o(const std::string& fname, const State& state, const BearingObservationVector& bearings,
               const OdometryObservationVector& odometries, int fixed_pose_id, bool with_landmarks) {
    FILE* f = std::fopen(fname.c_str(), "w");
    if (!f) return false;
    if (with_landmarks)
        for (int j = 0; j < state.number_of_landmarks(); j++)
            std::fprintf(f, "VERTEX_XY %d %.9g %.9g\n", state.landmark_id_at(j), state.landmark_at(j).x(), state.landmark_at(j).y());
    for (int i = 0; i < state.number_of_poses(); i++) {
        const EPose p = t2v(state.pose_at(i));
        std::fprintf(f, "VERTEX_SE2 %d %.9g %.9g %.9g\n", state.pose_id_at(i), p.x(), p.y(), p.z());
    }
    if (fixed_pose_id >= 0) std::fprintf(f, "FIX %d\n", fixed_pose_id);
    for (const OdometryObservation& o : odometries) {
        const EPose z = o.get_transformation();
        const la::Mat3f om = o.get_omega();
        std::fprintf(f, "EDGE_SE2 %d %d %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", o.get_source_id(), o.get_dest_id(), z.x(), z.y(), z.z(),
                     om(0, 0), om(0, 1), om(0, 2), om(1, 1), om(1, 2), om(2, 2));
    }
    for (const BearingObservation& b : bearings)
        std::fprintf(f, "EDGE_BEARING_SE2_XY %d %d %.9g %.9g\n", b.get_pose_id(), b.get_lm_id(), b.get_bearing().angle(), b.get_omega());
    std::fclose(f);
    return true;
}

}  // namespace proj02

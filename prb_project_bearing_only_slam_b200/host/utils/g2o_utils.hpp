// g2o_utils.hpp -- the reference's loader (utils/g2o_utils.hpp:27-29) plus a writer (the reference has none).
//
// Line formats (data/README.txt, g2o wiki):
//   VERTEX_SE2 id x y theta | VERTEX_XY id x y | FIX id | EDGE_SE2 i j x y theta <upper triangle of omega, 6 numbers>
//   EDGE_BEARING_SE2_XY id_pose id_landmark bearing <ignored>
#pragma once

#include <string>

#include "../framework/observation.hpp"
#include "../framework/state.hpp"

namespace proj02 {

// legacy 5-argument form kept by the reference for its old tests (utils/g2o_utils.cpp:5-8)
void parse_g2o(std::string fname, State& state, BearingObservationVector& bearings, int& fixed_pose_id, float& bound);

// fixed_pose_id = -1 when the file has no FIX line; bound = max |coordinate| + 3; numbers are read with std::stof (float)
void parse_g2o(std::string fname, State& state, BearingObservationVector& bearings, OdometryObservationVector& odometries,
               int& fixed_pose_id, float& bound);

// extension: dump a state and its edges in the same format (9 significant digits: floats round-trip exactly)
bool write_g2o(const std::string& fname, const State& state, const BearingObservationVector& bearings,
               const OdometryObservationVector& odometries, int fixed_pose_id, bool with_landmarks = true);

}  // namespace proj02

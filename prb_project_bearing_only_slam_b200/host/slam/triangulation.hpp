// triangulation.hpp -- slam/triangulation.hpp:8 of the reference; the per-landmark least squares runs on the GPU (K8).
#pragma once

#include "../framework/observation.hpp"
#include "../framework/state.hpp"

namespace proj02 {

// Adds one landmark per distinct landmark id of `observations` to `state`, in ASCENDING id order (the reference's
// std::map iteration, slam/triangulation.cpp:65-74), at the least-squares intersection of its bearing rays
// (ColPivHouseholderQR semantics incl. the rank-1 basic solution of a single observation, which also prints the
// reference's warning).  Unknown pose ids throw std::out_of_range.  device < 0 keeps the default device (0).
void triangulate_landmarks(State& state, const BearingObservationVector& observations);
void triangulate_landmarks(State& state, const BearingObservationVector& observations, int device, bool fp32);

}  // namespace proj02

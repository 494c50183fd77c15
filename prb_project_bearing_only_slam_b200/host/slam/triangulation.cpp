// triangulation.cpp -- host side of triangulate_landmarks: id bookkeeping here, arithmetic in bos_triangulate_landmarks.
#include "triangulation.hpp"

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/bos_b200.h"

namespace proj02 {

void triangulate_landmarks(State& state, const BearingObservationVector& observations) { triangulate_landmarks(state, observations, 0, false); }

void triangulate_landmarks(State& state, const BearingObservationVector& observations, int device, bool fp32) {
    // slam/triangulation.cpp:5-19: bucket by landmark id; std::map order = ascending id = the stix the landmarks get
    std::map<int, int> lm_index;
    std::map<int, int> count;
    for (const BearingObservation& o : observations) { lm_index[o.get_lm_id()] = 0; count[o.get_lm_id()]++; }
    int NL = 0;
    for (auto& kv : lm_index) kv.second = NL++;
    if (NL == 0) return;
    const int NP = state.number_of_poses();
    std::vector<double> poses(4 * (size_t)NP), lms(2 * (size_t)NL), z(observations.size());
    std::vector<int32_t> bp(observations.size()), bl(observations.size());
    for (int i = 0; i < NP; i++) {
        const NEPose& X = state.pose_at(i);
        poses[4 * (size_t)i] = X.translation().x(); poses[4 * (size_t)i + 1] = X.translation().y();
        poses[4 * (size_t)i + 2] = X.linear()(0, 0); poses[4 * (size_t)i + 3] = X.linear()(1, 0);
    }
    for (size_t e = 0; e < observations.size(); e++) {
        bp[e] = state.pose_stix(observations[e].get_pose_id());      // throws like get_pose_by_id in the reference
        bl[e] = lm_index[observations[e].get_lm_id()];
        z[e] = observations[e].get_bearing().angle();
    }
    for (const auto& kv : count)
        if (kv.second == 1) {   // slam/triangulation.cpp:38-42
            std::cout << "Landmark no. " << kv.first << " only has one observation.\n";
            std::cout << "  Bearing-only SLAM won't be able to locate it properly." << std::endl;
        }
    bos_options opt;
    bos_default_options(&opt);
    if (device >= 0) opt.device = device;
    opt.precision = fp32 ? BOS_PRECISION_F32 : BOS_PRECISION_F64;
    int single = 0;
    const int rc = bos_triangulate_landmarks(&opt, NP, poses.data(), (int64_t)observations.size(), bp.data(), bl.data(), z.data(), NL, lms.data(), &single);
    if (rc != BOS_OK) throw std::runtime_error("triangulate_landmarks: bos_triangulate_landmarks failed with status " + std::to_string(rc) +
                                               " (no CUDA device? there is no CPU fallback)");
    for (const auto& kv : lm_index) state.add_landmark(LMPos((float)lms[2 * (size_t)kv.second], (float)lms[2 * (size_t)kv.second + 1]), kv.first);
}

}  // namespace proj02

// state.hpp -- State with the reference's public surface (framework/state.hpp:15-54).
// Extensions (marked) give the GPU-backed Solver bulk access by state index; draw() is not provided (OpenCV UI).
#pragma once

#include "definitions.hpp"

namespace proj02 {

// framework/state.hpp:11-13 : LEFT perturbation
inline NEPose boxplus(const NEPose& X, const EPose& delta_x) { return v2t(delta_x) * X; }

class State {
 public:
    State(int expected_states = 300, int expected_landmarks = 200);

    void add_pose(const NEPose& pose, const int& id);
    void add_pose(const float& x, const float& y, const float& theta, const int& id);
    void add_landmark(const LMPos& lm, const int& id);
    void add_landmark(const float& x, const float& y, const int& id);

    NEPose get_pose_by_id(const int& id) const;       // throws std::out_of_range for an unknown id (std::map::at)
    LMPos get_landmark_by_id(const int& id) const;

    int number_of_poses() const;
    int number_of_landmarks() const;

    int pose_stix(const int& id) const;
    int landmark_stix(const int& id) const;

    int default_pose_id();

    // host-side convenience with the reference's semantics (framework/state.cpp:69-80); Solver::step() does NOT use it:
    // there the update runs on the GPU (K7) and the result is mirrored back into this object.
    void apply_boxplus(const la::VectorXf& delta_x);

    void print_full_vector();

    // ---- extensions used by the GPU-backed Solver and the harness -------------------------------------------------
    const NEPose& pose_at(int stix) const { return poses[stix]; }
    const LMPos& landmark_at(int stix) const { return landmarks[stix]; }
    int pose_id_at(int stix) const { return pose_stix_to_id[stix]; }
    int landmark_id_at(int stix) const { return lm_stix_to_id[stix]; }
    void set_pose_at(int stix, const NEPose& p) { poses[stix] = p; ++version_; }
    void set_landmark_at(int stix, const LMPos& l) { landmarks[stix] = l; ++version_; }
    // bumped by every mutation; the Solver re-uploads its device copy when it sees a version it did not write
    unsigned long long version() const { return version_; }

 private:
    NEPoseVector poses;
    LMPosVector landmarks;
    AssociationMap pose_id_to_stix;
    AssociationVec pose_stix_to_id;
    AssociationMap lm_id_to_stix;
    AssociationVec lm_stix_to_id;
    unsigned long long version_ = 0;
};

}  // namespace proj02

// state.hpp -- the estimate: poses and landmarks addressed by caller ids, stored by state index ("stix").  Public surface of the reference's
// State (framework/state.hpp:15-54) plus bulk access by stix for the GPU-backed Solver; draw() is not provided (OpenCV UI, out of scope).
#pragma once

#include "definitions.hpp"

namespace proj02 {

// id <-> stix bookkeeping, once for poses and once for landmarks.  stix = insertion order.  Inserting an id a second time re-points the id at
// the new stix while the earlier entry stays in the vectors (std::map overwrite, framework/state.cpp:20-41; tests/state_test.cpp:17-18).
class IdTable {
    std::map<int, int> stix_of_;
    std::vector<int> id_of_;

 public:
    void reserve(int n) { if (n > 0) id_of_.reserve((std::size_t)n); }
    int append(int id) {
        const int stix = (int)id_of_.size();
        stix_of_[id] = stix;
        id_of_.push_back(id);
        return stix;
    }
    int stix(int id) const { return stix_of_.at(id); }      // std::out_of_range for an unknown id, like the reference's map::at
    int id(int stix) const { return id_of_.at((std::size_t)stix); }
    int size() const { return (int)id_of_.size(); }
};

// LEFT perturbation (framework/state.hpp:11-13): the increment is applied in the world frame
inline NEPose boxplus(const NEPose& X, const EPose& delta_x) { return v2t(delta_x) * X; }

class State {
    NEPoseVector poses_;
    LMPosVector landmarks_;
    IdTable pose_ids_, lm_ids_;
    unsigned long long version_ = 0;   // bumped by every mutation: the Solver re-uploads its device copy when it sees a version it did not write

 public:
    State(int expected_states = 300, int expected_landmarks = 200);

    // ---- the reference's surface --------------------------------------------------------------------------------
    void add_pose(const float& x, const float& y, const float& theta, const int& id);
    void add_pose(const NEPose& pose, const int& id);
    void add_landmark(const float& x, const float& y, const int& id);
    void add_landmark(const LMPos& lm, const int& id);

    int number_of_poses() const { return (int)poses_.size(); }
    int number_of_landmarks() const { return (int)landmarks_.size(); }
    int pose_stix(const int& id) const { return pose_ids_.stix(id); }
    int landmark_stix(const int& id) const { return lm_ids_.stix(id); }
    NEPose get_pose_by_id(const int& id) const { return poses_[(std::size_t)pose_ids_.stix(id)]; }
    LMPos get_landmark_by_id(const int& id) const { return landmarks_[(std::size_t)lm_ids_.stix(id)]; }
    int default_pose_id() { return pose_ids_.id(0); }

    // host-side convenience with the reference's semantics (framework/state.cpp:69-80).  Solver::step() does NOT use it: there the
    // update runs on the GPU (K7) and the result is mirrored back into this object.
    void apply_boxplus(const la::VectorXf& delta_x);
    void print_full_vector();

    // ---- by state index (the Solver, the harness) ------------------------------------------------------------------
    const NEPose& pose_at(int stix) const { return poses_[(std::size_t)stix]; }
    const LMPos& landmark_at(int stix) const { return landmarks_[(std::size_t)stix]; }
    int pose_id_at(int stix) const { return pose_ids_.id(stix); }
    int landmark_id_at(int stix) const { return lm_ids_.id(stix); }
    void set_pose_at(int stix, const NEPose& p) { poses_[(std::size_t)stix] = p; ++version_; }
    void set_landmark_at(int stix, const LMPos& l) { landmarks_[(std::size_t)stix] = l; ++version_; }
    unsigned long long version() const { return version_; }
};

}  // namespace proj02

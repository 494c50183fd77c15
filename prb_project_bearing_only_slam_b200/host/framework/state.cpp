// state.cpp -- the out-of-line part of State (see state.hpp; semantics of framework/state.cpp:7-93 of the reference).
#include "state.hpp"

namespace proj02 {

State::State(int expected_states, int expected_landmarks) {
    if (expected_states > 0) poses_.reserve((std::size_t)expected_states);
    if (expected_landmarks > 0) landmarks_.reserve((std::size_t)expected_landmarks);
    pose_ids_.reserve(expected_states);
    lm_ids_.reserve(expected_landmarks);
}

void State::add_pose(const NEPose& pose, const int& id) {
    pose_ids_.append(id);
    poses_.push_back(pose);
    ++version_;
}
void State::add_pose(const float& x, const float& y, const float& theta, const int& id) { add_pose(v2t(EPose(x, y, theta)), id); }

void State::add_landmark(const LMPos& lm, const int& id) {
    lm_ids_.append(id);
    landmarks_.push_back(lm);
    ++version_;
}
void State::add_landmark(const float& x, const float& y, const int& id) { add_landmark(LMPos(x, y), id); }

void State::apply_boxplus(const la::VectorXf& delta_x) {
    const std::size_t np = poses_.size();
    for (std::size_t i = 0; i < np; i++) poses_[i] = boxplus(poses_[i], delta_x.segment<3>(3 * i));
    for (std::size_t j = 0; j < landmarks_.size(); j++) landmarks_[j] += delta_x.segment<2>(3 * np + 2 * j);   // landmarks are Euclidean
    ++version_;
}

void State::print_full_vector() {
    std::cout << "State:";
    for (const NEPose& p : poses_) std::cout << " " << t2v(p);
    for (const LMPos& l : landmarks_) std::cout << " " << l;
    std::cout << std::endl;
}

}  // namespace proj02

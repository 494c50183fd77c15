// state.cpp -- framework/state.cpp:7-93 of the reference, same semantics: stix = insertion order; a duplicate id
// overwrites the id->stix map entry while both vectors keep growing (tests/state_test.cpp:17-18).
#include "state.hpp"

namespace proj02 {

State::State(int expected_states, int expected_landmarks) {
    poses.reserve(expected_states > 0 ? expected_states : 0);
    landmarks.reserve(expected_landmarks > 0 ? expected_landmarks : 0);
    pose_stix_to_id.reserve(expected_states > 0 ? expected_states : 0);
    lm_stix_to_id.reserve(expected_landmarks > 0 ? expected_landmarks : 0);
}

void State::add_pose(const NEPose& pose, const int& id) {
    pose_id_to_stix[id] = (int)poses.size();
    poses.push_back(pose);
    pose_stix_to_id.push_back(id);
    ++version_;
}
void State::add_pose(const float& x, const float& y, const float& theta, const int& id) { add_pose(v2t(EPose(x, y, theta)), id); }

void State::add_landmark(const LMPos& lm, const int& id) {
    lm_id_to_stix[id] = (int)landmarks.size();
    landmarks.push_back(lm);
    lm_stix_to_id.push_back(id);
    ++version_;
}
void State::add_landmark(const float& x, const float& y, const int& id) { add_landmark(LMPos(x, y), id); }

NEPose State::get_pose_by_id(const int& id) const { return poses[pose_id_to_stix.at(id)]; }
LMPos State::get_landmark_by_id(const int& id) const { return landmarks[lm_id_to_stix.at(id)]; }
int State::number_of_poses() const { return (int)poses.size(); }
int State::number_of_landmarks() const { return (int)landmarks.size(); }
int State::pose_stix(const int& id) const { return pose_id_to_stix.at(id); }
int State::landmark_stix(const int& id) const { return lm_id_to_stix.at(id); }
int State::default_pose_id() { return pose_stix_to_id.at(0); }

void State::apply_boxplus(const la::VectorXf& delta_x) {
    const int NP = (int)poses.size(), NL = (int)landmarks.size();
    for (int i = 0; i < NP; i++) poses[i] = boxplus(poses[i], delta_x.segment<3>(3 * (std::size_t)i));
    for (int j = 0; j < NL; j++) landmarks[j] += delta_x.segment<2>(3 * (std::size_t)NP + 2 * (std::size_t)j);
    ++version_;
}

void State::print_full_vector() {
    std::cout << "State:";
    for (const NEPose& p : poses) std::cout << " " << t2v(p);
    for (const LMPos& l : landmarks) std::cout << " " << l;
    std::cout << std::endl;
}

}  // namespace proj02

// observation.hpp -- the two edge types of the problem with the reference's public surface (framework/observation.hpp:12-87:
// constructors, get_* accessors, the three container typedefs), so that code written against the reference compiles unchanged.
// Storage is plain numbers (ids, one angle / a 3-vector, the information values): the Solver copies them ONCE into the device's SoA edge
// buffers, nothing here is touched per iteration.
#pragma once

#include <array>

#include "definitions.hpp"

namespace proj02 {

// One bearing measurement of a landmark taken from a pose.  The angle is kept exactly as given (the reference normalises it
// on every use through Rotation2D::smallestAngle, so does the device kernel); the information value defaults to 1.
class BearingObservation {
    int ends_[2];     // pose id, landmark id
    float z_;
    float info_;

 public:
    BearingObservation(const int& pose_id, const int& lm_id, const float& bearing, const float& omega = 1) : ends_{pose_id, lm_id}, z_(bearing), info_(omega) {}
    BearingObservation(const int& pose_id, const int& lm_id, const Rotation2f& bearing, const float& omega = 1)
        : BearingObservation(pose_id, lm_id, bearing.angle(), omega) {}

    int get_pose_id() const { return ends_[0]; }
    int get_lm_id() const { return ends_[1]; }
    float get_omega() const { return info_; }
    Rotation2f get_bearing() const { return Rotation2f(z_); }
};

// One odometry measurement between two poses, expressed on the chart of the SOURCE pose:
//   z = [ R_s^T (t_d - t_s) ; theta_d - theta_s ]   (not a homogeneous transform; the g2o EDGE_SE2 convention)
// with a full 3x3 information matrix (row-major here).
class OdometryObservation {
    int ends_[2];                 // source id, destination id
    std::array<float, 3> z_;
    std::array<float, 9> info_;

    void keep(const la::Mat3f& m) {
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) info_[3 * r + c] = m(r, c);
    }

 public:
    OdometryObservation(const int& source_id, const int& dest_id, float x, float y, float theta, la::Mat3f omega) : ends_{source_id, dest_id}, z_{{x, y, theta}} {
        keep(omega);
    }
    OdometryObservation(const int& source_id, const int& dest_id, EPose transformation, la::Mat3f omega)
        : OdometryObservation(source_id, dest_id, transformation(0), transformation(1), transformation(2), omega) {}

    int get_source_id() const { return ends_[0]; }
    int get_dest_id() const { return ends_[1]; }
    EPose get_transformation() const { return EPose(z_[0], z_[1], z_[2]); }
    la::Mat3f get_omega() const {
        la::Mat3f m;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) m(r, c) = info_[3 * r + c];
        return m;
    }
    // all nine entries as structural non-zeros, explicit zeros included (framework/observation.cpp:34-51)
    SparseMatrixXf get_omega_sparse() const {
        SparseMatrixXf m(3, 3);
        for (int k = 0; k < 9; k++) m.coeffRef(k / 3, k % 3) = info_[k];
        return m;
    }
    // bulk access for the SoA conversion
    const std::array<float, 3>& measurement() const { return z_; }
    const std::array<float, 9>& information() const { return info_; }
};

using BearingObservationVector = std::vector<BearingObservation>;
using OdometryObservationVector = std::vector<OdometryObservation>;
using BearingObservationsByLandmarkId = std::map<int, BearingObservationVector>;

}  // namespace proj02

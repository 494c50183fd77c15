// observation.hpp -- the reference's edge types (framework/observation.hpp:12-87), host AoS; converted once to the
// device SoA edge buffers by the Solver.
#pragma once

#include "definitions.hpp"

namespace proj02 {

// bearing-only pose-landmark observation; the bearing is stored un-normalised, omega defaults to 1
class BearingObservation {
 public:
    BearingObservation(const int& pose_id, const int& lm_id, const Rotation2f& bearing, const float& omega = 1)
        : pose_id(pose_id), lm_id(lm_id), bearing(bearing), omega(omega) {}
    BearingObservation(const int& pose_id, const int& lm_id, const float& bearing, const float& omega = 1)
        : pose_id(pose_id), lm_id(lm_id), bearing(bearing), omega(omega) {}
    int get_pose_id() const { return pose_id; }
    int get_lm_id() const { return lm_id; }
    Rotation2f get_bearing() const { return bearing; }
    float get_omega() const { return omega; }

 private:
    int pose_id;
    int lm_id;
    Rotation2f bearing;
    float omega;
};

// odometry on the chart of the source pose: z = [R_s^T (t_d - t_s) ; theta_d - theta_s]
class OdometryObservation {
 public:
    OdometryObservation(const int& source_id, const int& dest_id, EPose transformation, la::Mat3f omega)
        : source_id(source_id), dest_id(dest_id), transformation(transformation), omega(omega) {}
    OdometryObservation(const int& source_id, const int& dest_id, float x, float y, float theta, la::Mat3f omega)
        : source_id(source_id), dest_id(dest_id), transformation(x, y, theta), omega(omega) {}
    int get_source_id() const { return source_id; }
    int get_dest_id() const { return dest_id; }
    EPose get_transformation() const { return transformation; }
    la::Mat3f get_omega() const { return omega; }
    SparseMatrixXf get_omega_sparse() const {   // framework/observation.cpp:34-51: all nine entries, explicit zeros kept
        SparseMatrixXf m(3, 3);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) m.coeffRef(i, j) = omega(i, j);
        return m;
    }

 private:
    int source_id;
    int dest_id;
    EPose transformation;
    la::Mat3f omega;
};

typedef std::vector<BearingObservation> BearingObservationVector;
typedef std::vector<OdometryObservation> OdometryObservationVector;
typedef std::map<int, BearingObservationVector> BearingObservationsByLandmarkId;

}  // namespace proj02

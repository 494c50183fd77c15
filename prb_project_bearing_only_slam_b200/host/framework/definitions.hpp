// definitions.hpp -- the reference's typedefs and t2v / v2t (framework/definitions.hpp:15-53) over linalg.hpp.
// The OpenCV image type of the reference (RGBImage) is not part of this host: drawing is out of scope.
#pragma once

#include <map>
#include <vector>

#include "linalg.hpp"

namespace proj02 {

typedef la::Iso2f NEPose;    // Non-Euclidean Pose, a 2D homogeneous transform
typedef la::Vec3f EPose;     // Euclidean Pose (x, y, theta)
typedef la::Vec2f LMPos;     // Landmark Position

typedef std::vector<NEPose> NEPoseVector;
typedef std::vector<LMPos> LMPosVector;

typedef la::Rotation2f Rotation2f;

typedef std::map<int, int> AssociationMap;   // id -> state index
typedef std::vector<int> AssociationVec;     // state index -> id

typedef la::SparseMatrixXf SparseMatrixXf;
typedef la::Triplet_f Triplet_f;

// framework/definitions.hpp:39-43
inline EPose t2v(const NEPose& nep) {
    const la::Vec2f t = nep.translation();
    const Rotation2f r = Rotation2f(nep.rotation());
    return EPose(t.x(), t.y(), r.smallestAngle());
}

// framework/definitions.hpp:45-53
inline NEPose v2t(const EPose& ep) {
    NEPose X;
    X.setIdentity();
    X.translation() = la::Vec2f(ep.x(), ep.y());
    X.linear() = Rotation2f(ep.z()).matrix();
    return X;
}

}  // namespace proj02

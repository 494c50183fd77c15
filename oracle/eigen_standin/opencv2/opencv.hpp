// Stand-in for <opencv2/opencv.hpp> -- TEST INFRASTRUCTURE ONLY (see ../Eigen/standin.hpp).
// The reference's hot path takes two things from OpenCV: the RGBImage typedef (framework/definitions.hpp:16, drawing only) and the
// constants CV_PI / CV_2PI used by Solver::normalized_angle (slam/solver_jacobians.cpp:325-333).  The constants below are the values
// OpenCV's core/cvdef.h publishes (double literals).
#pragma once
// OpenCV's headers pull in the standard headers the reference then uses without including them itself (framework/definitions.hpp)
#include <algorithm>
#include <cmath>
#include <iostream>
#include <map>
#include <math.h>
#include <string>
#include <vector>
#define CV_PI 3.1415926535897932384626433832795
#define CV_2PI 6.283185307179586476925286766559
namespace cv {
struct Vec3b { unsigned char v[3]; };
template <class T> class Mat_ {};
}  // namespace cv

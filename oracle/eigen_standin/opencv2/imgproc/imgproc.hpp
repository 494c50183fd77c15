// Stand-in for <opencv2/imgproc/imgproc.hpp> (declarations of utils/draw_utils.hpp only need the types of opencv.hpp).
#pragma once
#include "../opencv.hpp"

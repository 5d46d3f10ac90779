// TEST STUB — not OpenCV.  See tests/cpp/stubs/Eigen/Eigen.
#pragma once
#include <string>
#include "../core/core.hpp"
namespace cv {
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }
}  // namespace cv

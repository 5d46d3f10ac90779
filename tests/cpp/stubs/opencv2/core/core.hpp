// TEST STUB — not OpenCV.  See tests/cpp/stubs/Eigen/Eigen.
#pragma once
#include <cstddef>
#define CV_8UC1 0
namespace cv {
class Mat {
 public:
  Mat() : data(nullptr), rows(0), cols(0), step(0) {}
  Mat(int r, int c, int /*type*/, void* d, size_t s = 0) : data(static_cast<unsigned char*>(d)), rows(r), cols(c), step(s ? s : (size_t)c) {}
  unsigned char* data;
  int rows, cols;
  size_t step;   // cv::MatStep converts to size_t
};
}  // namespace cv

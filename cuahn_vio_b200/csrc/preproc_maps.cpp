// Host-side generation of the undistort + resize maps, Eigen/OpenCV-free.
// Reference: CamBase::initialize_undist_map[_fisheye] (cuahn_ros/ov_core/src/cam/CamBase.h:165-180), i.e.
// cv::initUndistortRectifyMap / cv::fisheye::initUndistortRectifyMap with R = I, the fixed 90-degree target camera and
// CV_32FC1 output.  Every operation is double precision in the library's order (running sums along a row included), so
// the float maps are bit-identical to OpenCV's (tests/test_preproc.py against OpenCV-made fixtures).
#include <cmath>

#include "../../include/uahn_preproc.h"

extern "C" int uahn_undistort_init_maps(int fisheye, const double* k, const double* d, float* map1, float* map2) {
  if (!k || !d || !map1 || !map2) return UAHN_ERR_INVALID;
  const double fov = 45.0 * 2.0, pi = 2.0 * std::acos(0.0);                       // CamBase.h:166-167
  const double nf = (320.0 - 1.0) / 2.0 / std::tan(fov / 180.0 * pi / 2.0);       // :169
  const double ncx = (320.0 - 1.0) / 2.0, ncy = (224.0 - 1.0) / 2.0;
  // ir = (newK * R)^-1, R = I
  const double ir[9] = {1.0 / nf, 0.0, -ncx / nf, 0.0, 1.0 / nf, -ncy / nf, 0.0, 0.0, 1.0};
  const double fx = k[0], fy = k[1], cx = k[2], cy = k[3];
  for (int i = 0; i < UAHN_IMG_H; ++i) {
    double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
    for (int j = 0; j < UAHN_IMG_W; ++j, _x += ir[0], _y += ir[3], _w += ir[6]) {
      double u, v;
      if (fisheye) {
        const double x = _x / _w, y = _y / _w;
        const double r = std::sqrt(x * x + y * y);
        const double theta = std::atan(r);
        const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
        const double theta_d = theta * (1 + d[0] * t2 + d[1] * t4 + d[2] * t6 + d[3] * t8);
        const double scale = (r == 0) ? 1.0 : theta_d / r;
        u = fx * x * scale + cx;
        v = fy * y * scale + cy;
      } else {
        const double w = 1.0 / _w, x = _x * w, y = _y * w;
        const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
        const double kr = 1 + ((0.0 * r2 + d[1]) * r2 + d[0]) * r2;               // k3..k6 = 0 in the 4-coefficient model
        u = fx * (x * kr + d[2] * _2xy + d[3] * (r2 + 2 * x2)) + cx;
        v = fy * (y * kr + d[2] * (r2 + 2 * y2) + d[3] * _2xy) + cy;
      }
      map1[i * UAHN_IMG_W + j] = (float)u;
      map2[i * UAHN_IMG_W + j] = (float)v;
    }
  }
  return UAHN_OK;
}

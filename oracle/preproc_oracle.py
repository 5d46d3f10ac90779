"""TEST INFRASTRUCTURE — numpy restatement of the pre-processing step in front of the UAHN path.

Only tests/ may import this.  The reference step is
  CamBase::initialize_undist_map[_fisheye]   /root/reference/cuahn_ros/ov_core/src/cam/CamBase.h:165-180
  CamBase::undistort_and_resize_img          /root/reference/cuahn_ros/ov_core/src/cam/CamBase.h:182-186
whose arithmetic lives in a third-party dependency that is not under /root/reference: OpenCV (the ROS system
package; the reference pins no version) — cv::initUndistortRectifyMap, cv::fisheye::initUndistortRectifyMap and
cv::remap(INTER_LINEAR, BORDER_CONSTANT 0).  Their published algorithms are restated below.  Pinning: fixtures in
tests/golden/preproc.npz were produced by OpenCV 4.13 itself in the build container (oracle/make_golden_preproc.py);
tests/test_preproc.py requires this restatement to reproduce cv::remap bit for bit and the maps to float rounding.
"""
from __future__ import annotations

import numpy as np

OUT_H, OUT_W = 224, 320
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS            # cv::remap works on 1/32-pixel fixed-point coordinates
INTER_REMAP_COEF_BITS = 15


def standard_camera():
    """The 90-degree-FoV pinhole the network was trained on (CamBase.h:166-169): f = 159.5, c = (159.5, 111.5)."""
    fov = 45.0 * 2.0
    pi = 2.0 * np.arccos(0.0)
    f = (320.0 - 1.0) / 2.0 / np.tan(fov / 180.0 * pi / 2.0)
    return f, f, (320.0 - 1.0) / 2.0, (224.0 - 1.0) / 2.0


def init_undistort_maps(fisheye: bool, k, d):
    """map1 (x), map2 (y) float32 [224, 320]: raw-image coordinates sampled by every output pixel.

    k = (fx, fy, cx, cy) and d = 4 distortion coefficients of the raw camera; R = I; new camera = standard_camera().
    radtan : cv::initUndistortRectifyMap, d = (k1, k2, p1, p2)
    fisheye: cv::fisheye::initUndistortRectifyMap (equidistant), d = (k1..k4)
    Both walk each row with running sums (_x += ir[0] per column), reproduced here with np.cumsum-free exact replay.
    """
    fx, fy, cx, cy = [float(v) for v in k]
    nfx, nfy, ncx, ncy = standard_camera()
    ir = np.array([[1.0 / nfx, 0.0, -ncx / nfx], [0.0, 1.0 / nfy, -ncy / nfy], [0.0, 0.0, 1.0]])
    m1 = np.empty((OUT_H, OUT_W), np.float32)
    m2 = np.empty((OUT_H, OUT_W), np.float32)
    for i in range(OUT_H):
        # running sums exactly as the library accumulates them along a row
        _x = np.empty(OUT_W); _y = np.empty(OUT_W); _w = np.empty(OUT_W)
        ax, ay, aw = i * ir[0, 1] + ir[0, 2], i * ir[1, 1] + ir[1, 2], i * ir[2, 1] + ir[2, 2]
        for j in range(OUT_W):
            _x[j], _y[j], _w[j] = ax, ay, aw
            ax += ir[0, 0]; ay += ir[1, 0]; aw += ir[2, 0]
        if fisheye:
            x, y = _x / _w, _y / _w
            r = np.sqrt(x * x + y * y)
            theta = np.arctan(r)
            t2 = theta * theta; t4 = t2 * t2; t6 = t4 * t2; t8 = t4 * t4
            theta_d = theta * (1 + d[0] * t2 + d[1] * t4 + d[2] * t6 + d[3] * t8)
            scale = np.where(r == 0, 1.0, theta_d / np.where(r == 0, 1.0, r))
            u, v = fx * x * scale + cx, fy * y * scale + cy
        else:
            w = 1.0 / _w
            x, y = _x * w, _y * w
            x2, y2 = x * x, y * y
            r2, _2xy = x2 + y2, 2 * x * y
            kr = 1 + ((0.0 * r2 + d[1]) * r2 + d[0]) * r2          # k3 = 0, denominator 1 (4-coefficient model)
            u = fx * (x * kr + d[2] * _2xy + d[3] * (r2 + 2 * x2)) + cx
            v = fy * (y * kr + d[2] * (r2 + 2 * y2) + d[3] * _2xy) + cy
        m1[i], m2[i] = u.astype(np.float32), v.astype(np.float32)
    return m1, m2


def remap_bilinear_u8(src: np.ndarray, map1: np.ndarray, map2: np.ndarray) -> np.ndarray:
    """cv::remap(src u8, CV_32FC1 maps, INTER_LINEAR, BORDER_CONSTANT, 0).

    Coordinates are rounded (half to even, cvRound) to 1/32 pixel; the four taps are blended with the integer table
    weights (32 - fx)(32 - fy) * 32 ... (they sum to 2^15 exactly for the bilinear table) and the result is
    (sum + 2^14) >> 15; taps outside the image contribute 0.
    """
    H, W = src.shape
    sx = np.rint(map1.astype(np.float32) * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    sy = np.rint(map2.astype(np.float32) * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    # XY is stored as saturate_cast<short>
    x0 = np.clip(sx >> INTER_BITS, -32768, 32767)
    y0 = np.clip(sy >> INTER_BITS, -32768, 32767)
    fx, fy = sx & (INTER_TAB_SIZE - 1), sy & (INTER_TAB_SIZE - 1)
    s = src.astype(np.int64)

    def tap(y, x):
        ok = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        return np.where(ok, s[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)], 0)
    w00 = (INTER_TAB_SIZE - fx) * (INTER_TAB_SIZE - fy) * 32
    w01 = fx * (INTER_TAB_SIZE - fy) * 32
    w10 = (INTER_TAB_SIZE - fx) * fy * 32
    w11 = fx * fy * 32
    acc = tap(y0, x0) * w00 + tap(y0, x0 + 1) * w01 + tap(y0 + 1, x0) * w10 + tap(y0 + 1, x0 + 1) * w11
    return ((acc + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS).astype(np.uint8)

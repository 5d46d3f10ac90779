"""(container only) Fixtures for the undistort + resize step from OpenCV itself -> tests/golden/preproc.npz.

    python -m oracle.make_golden_preproc

Stores, for the two shipped UZH-FPV camera configurations (uzhfpv.launch:81-82, 93-94; fisheye) and one radtan model:
the camera parameters, the positions/values where OpenCV's maps differ from the restatement's (normally a handful of
float-rounding cases), and cv::remap's output on a seeded raw frame.
"""
import os

import cv2
import numpy as np

from oracle import preproc_oracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAMS = {   # name: (fisheye, k, d)
    "uzhfpv_indoor_fwd": (True, [275.46015578667294, 274.9948095922592, 315.958384100568, 242.7123497822731],
                          [-6.545154718304953e-06, -0.010379525898159981, 0.014935312423953146, -0.005639061406567785]),
    "uzhfpv_outdoor_45": (True, [275.3385453506587, 275.0852058534152, 315.7697752181792, 233.72625444124952],
                          [-0.017811595366268803, 0.04897078939103475, -0.041363300782847834, 0.011440891936886532]),
    "radtan_test": (False, [458.654, 457.296, 327.127, 238.253], [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]),
}


def raw_frame(seed: int) -> np.ndarray:
    from cuahn_vio_b200.synthetic import synthetic_raw_frame
    return synthetic_raw_frame(seed)


def main():
    out = {}
    nfx, nfy, ncx, ncy = P.standard_camera()
    newK = np.array([[nfx, 0, ncx], [0, nfy, ncy], [0, 0, 1]])
    for name, (fisheye, k, d) in CAMS.items():
        K = np.array([[k[0], 0, k[2]], [0, k[1], k[3]], [0, 0, 1]])
        D = np.array(d, np.float64)
        if fisheye:
            m1, m2 = cv2.fisheye.initUndistortRectifyMap(K, D, np.eye(3), newK, (320, 224), cv2.CV_32FC1)
        else:
            m1, m2 = cv2.initUndistortRectifyMap(K, D, None, newK, (320, 224), cv2.CV_32FC1)
        o1, o2 = P.init_undistort_maps(fisheye, k, d)
        raw = raw_frame(7)
        ref = cv2.remap(raw, m1, m2, cv2.INTER_LINEAR)
        for tag, cvm, om in (("m1", m1, o1), ("m2", m2, o2)):
            idx = np.flatnonzero(cvm.ravel() != om.ravel())
            out[f"{name}_{tag}_diff_idx"] = idx.astype(np.int32)
            out[f"{name}_{tag}_diff_val"] = cvm.ravel()[idx]
            print(name, tag, "maps differ at", idx.size, "of", cvm.size, "max |d| =",
                  float(np.abs(cvm - om).max()))
        mine = P.remap_bilinear_u8(raw, m1, m2)
        print(name, "remap restatement vs cv2.remap: mismatches", int((mine != ref).sum()))
        out[f"{name}_fisheye"] = np.array(int(fisheye))
        out[f"{name}_k"] = np.array(k)
        out[f"{name}_d"] = np.array(d)
        out[f"{name}_remap"] = ref
        # a frame that also samples outside the raw image: shift the maps (border taps -> 0)
        ref_b = cv2.remap(raw, m1 + 200.0, m2 - 150.0, cv2.INTER_LINEAR)
        out[f"{name}_remap_shifted"] = ref_b
        print(name, "shifted: mismatches", int((P.remap_bilinear_u8(raw, m1 + np.float32(200.0), m2 - np.float32(150.0)) != ref_b).sum()))
    out["raw_seed"] = np.array(7)
    out["raw_checksum"] = np.array(int(raw_frame(7).astype(np.int64).sum()))
    out["opencv_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preproc.npz"), **out)


if __name__ == "__main__":
    main()

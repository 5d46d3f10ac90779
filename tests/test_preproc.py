"""SURVEY §8f row 2 — undistort + resize in front of the path (CamBase.h:165-186) against OpenCV-made fixtures.

CPU: the numpy restatement and the host map generator reproduce OpenCV's maps and cv::remap bit for bit.
GPU: the remap kernel, through the C ABI, reproduces cv::remap bit for bit and feeds the image ring."""
import os

import numpy as np
import pytest

from cuahn_vio_b200 import synthetic as S
from oracle import preproc_oracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAMS = ["uzhfpv_indoor_fwd", "uzhfpv_outdoor_45", "radtan_test"]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "preproc.npz"))


@pytest.fixture(scope="module")
def raw(golden):
    f = S.synthetic_raw_frame(int(golden["raw_seed"]))
    assert int(f.astype(np.int64).sum()) == int(golden["raw_checksum"])     # the frame the fixtures were made from
    return f


def _opencv_maps(golden, name, m1, m2):
    """OpenCV's maps = the restatement's, patched where the fixture recorded a difference (none at generation time)."""
    out = []
    for tag, m in (("m1", m1), ("m2", m2)):
        m = m.copy().ravel()
        m[golden[f"{name}_{tag}_diff_idx"]] = golden[f"{name}_{tag}_diff_val"]
        out.append(m.reshape(224, 320))
    return out


@pytest.fixture(scope="module")
def api():
    from cuahn_vio_b200 import build
    build.build()
    from cuahn_vio_b200 import api as a
    a.load_library()
    return a


@pytest.mark.parametrize("name", CAMS)
def test_restatement_and_host_maps_match_opencv(api, golden, raw, name):
    fisheye, k, d = bool(golden[f"{name}_fisheye"]), golden[f"{name}_k"], golden[f"{name}_d"]
    o1, o2 = P.init_undistort_maps(fisheye, k, d)
    assert golden[f"{name}_m1_diff_idx"].size == 0 and golden[f"{name}_m2_diff_idx"].size == 0   # restatement == OpenCV
    m1, m2 = api.undistort_init_maps(fisheye, k, d)               # C++ host code behind the ABI
    assert np.array_equal(m1, o1) and np.array_equal(m2, o2)      # bit-identical float maps
    c1, c2 = _opencv_maps(golden, name, o1, o2)
    assert np.array_equal(P.remap_bilinear_u8(raw, c1, c2), golden[f"{name}_remap"])
    assert np.array_equal(P.remap_bilinear_u8(raw, c1 + np.float32(200.0), c2 - np.float32(150.0)),
                          golden[f"{name}_remap_shifted"])
    # sanity of the geometry: the principal point of the raw camera lands on the centre of the target camera
    assert abs(m1[111:113, 159:161].mean() - k[2]) < 1.0 and abs(m2[111:113, 159:161].mean() - k[3]) < 1.0


def test_remap_identity_and_border():
    img = np.random.default_rng(0).integers(0, 256, (48, 64), dtype=np.uint8)
    v, u = np.meshgrid(np.arange(224, dtype=np.float32), np.arange(320, dtype=np.float32), indexing="ij")
    out = P.remap_bilinear_u8(img, u, v)
    assert np.array_equal(out[:48, :64], img) and out[48:, :].max() == 0 and out[:, 64:].max() == 0
    half = P.remap_bilinear_u8(img, u + np.float32(0.5), v)        # exact midpoints round like (a + b + 1) >> 1
    exp = (img[:, :-1].astype(np.int32) + img[:, 1:] + 1) >> 1
    assert np.array_equal(half[:48, :63], exp)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CAMS)
def test_gpu_remap_bit_exact(golden, raw, name):
    from cuahn_vio_b200 import api, weights
    wfile = weights.synthetic_weights_file(0)
    fisheye, k, d = bool(golden[f"{name}_fisheye"]), golden[f"{name}_k"], golden[f"{name}_d"]
    m1, m2 = api.undistort_init_maps(fisheye, k, d)
    c1, c2 = _opencv_maps(golden, name, m1, m2)
    with api.Uahn(wfile, "prior1", precision="fp32", max_batch=1) as net:
        with pytest.raises(api.UahnError):
            net.load_raw_image(raw, 0.0)                           # maps not set yet
        net.set_undistort_maps(480, 640, c1, c2)
        assert np.array_equal(net.stage_undistort(raw), golden[f"{name}_remap"])
        net.set_undistort_maps(480, 640, c1 + np.float32(200.0), c2 - np.float32(150.0))
        assert np.array_equal(net.stage_undistort(raw), golden[f"{name}_remap_shifted"])
        padded = np.zeros((480, 700), np.uint8); padded[:, :640] = raw
        assert np.array_equal(net.stage_undistort(padded[:, :640]), golden[f"{name}_remap_shifted"])   # row stride > cols
        with pytest.raises(api.UahnError):
            net.stage_undistort(raw[:100])


@pytest.mark.gpu
def test_raw_frames_feed_the_ring(golden, raw):
    """uahn_load_raw_image == undistort_and_resize_img + load_current_img (VioManager.cpp:181-188)."""
    from cuahn_vio_b200 import api, weights
    wfile = weights.synthetic_weights_file(0)
    name = CAMS[0]
    m1, m2 = api.undistort_init_maps(True, golden[f"{name}_k"], golden[f"{name}_d"])
    raw2 = np.roll(raw, 5, axis=1)
    f1, f2 = P.remap_bilinear_u8(raw, m1, m2), P.remap_bilinear_u8(raw2, m1, m2)
    prior = np.zeros(8)
    with api.Uahn(wfile, "prior3", precision="fp32", max_batch=1) as a, \
            api.Uahn(wfile, "prior3", precision="fp32", max_batch=1) as b:
        a.set_undistort_maps(480, 640, m1, m2)
        a.load_raw_image(raw, 1.0)
        a.load_raw_image(raw2, 2.0)
        b.load_image(f1, 1.0)
        b.load_image(f2, 2.0)
        assert a.img_counter == 2 and a.latest_inference_time == 2.0
        ma, ca, _ = a.infer(prior, seed=1, pair_index=0)
        mb, cb, _ = b.infer(prior, seed=1, pair_index=0)
        assert np.array_equal(ma, mb) and np.array_equal(ca, cb)

"""The C++ drop-in class (include/HomographyNet.h) compiles here; on the GPU box it must agree with the C ABI."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    from cuahn_vio_b200 import build
    lib = build.build()
    exe = str(tmp_path / "shim_main")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "shim_main.cpp"), "-o", exe, "-L", os.path.dirname(lib), "-luahn",
           "-Wl,-rpath," + os.path.dirname(lib)]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def test_shim_compiles_and_links(tmp_path):
    assert os.path.exists(_compile(tmp_path))


@pytest.mark.gpu
def test_shim_matches_c_abi(tmp_path):
    from cuahn_vio_b200 import api, synthetic as S, weights
    exe = _compile(tmp_path)
    wfile = weights.synthetic_weights_file(0)
    prev, curr, _, prior = S.synthetic_batch(2, start=40)
    frames = np.stack([prev[0], curr[0], curr[1]])
    pri = np.stack([prior[0], prior[0], prior[1]]).reshape(3, 8).astype(np.float64)
    frames.tofile(tmp_path / "frames.u8")
    pri.tofile(tmp_path / "priors.f64")
    out = subprocess.run([exe, wfile, str(tmp_path / "frames.u8"), "3", str(tmp_path / "priors.f64"), "0"],
                         check=True, capture_output=True, text=True).stdout
    assert "HNet cannot inference! Only has one image!" in out        # HomographyNet.cpp:155-158
    res = [l.split() for l in out.splitlines() if l.startswith("RESULT")]
    assert len(res) == 3 and all(float(v) == 0 for v in res[0][4:12])  # first frame: outputs untouched
    assert [int(r[2]) for r in res] == [1, 2, 3] and float(res[1][3]) == 11.0
    with api.Uahn(wfile, "prior3", precision="fp32", max_batch=1) as net:
        # the shim numbers inferences 0,1,... and seeds Philox with (seed=9, pair_index=inference count)
        m1, c1, _ = net.infer_batch(frames[0:1], frames[1:2], pri[1:2].astype(np.float32), seed=9, first_pair=0)
        m2, c2, _ = net.infer_batch(frames[1:2], frames[2:3], pri[2:3].astype(np.float32), seed=9, first_pair=1)
    for r, m, c in ((res[1], m1, c1), (res[2], m2, c2)):
        assert np.allclose(np.array(r[4:12], float), m[0], atol=1e-6)
        assert np.allclose(np.array(r[12:20], float), np.diag(c[0]), rtol=1e-6)


def test_reference_signature_flavour_type_checks():
    """`-DUAHN_WITH_EIGEN_OPENCV` (cv::Mat / Eigen::Matrix signatures of HomographyNet.h:26-32) against VioManager's call
    pattern, with the minimal stub headers under tests/cpp/stubs: the container has neither Eigen nor OpenCV, so this
    is a syntax + type check (`-fsyntax-only`), not a link."""
    cmd = ["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-Werror", "-DUAHN_WITH_EIGEN_OPENCV",
           "-I", os.path.join(ROOT, "tests", "cpp", "stubs"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "shim_reference_signatures.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_shim_iterative_slot_runs_the_graph_the_file_holds(tmp_path):
    """HomographyNet.cpp:104-124,209-230: the second model slot runs whatever graph its file holds.  A flat file exported
    from a TorchScript trace records the variant (weights.export_torchscript); here the records are written directly for a
    1-block and a 2-block iterative model, and the shim's iteration 1 must equal a PRIOR1 / PRIOR2 handle."""
    from cuahn_vio_b200 import api, synthetic as S, weights
    exe = _compile(tmp_path)
    wfile = weights.synthetic_weights_file(0)
    sd = S.synthetic_state_dict(0)
    prev, curr, _, prior = S.synthetic_batch(1, start=41)
    frames = np.stack([prev[0], curr[0]])
    pri = np.stack([prior[0], prior[0]]).reshape(2, 8).astype(np.float64)
    frames.tofile(tmp_path / "frames.u8")
    pri.tofile(tmp_path / "priors.f64")
    for variant in ("prior1", "prior2"):
        it_file = str(tmp_path / f"iter_{variant}.bin")
        weights.export_state_dict(sd, it_file, meta={"variant": weights.VARIANT_IDS[variant], "show_error": 0})
        out = subprocess.run([exe, wfile, str(tmp_path / "frames.u8"), "2", str(tmp_path / "priors.f64"), "0", it_file],
                             check=True, capture_output=True, text=True).stdout
        it = [l.split() for l in out.splitlines() if l.startswith("ITER")]
        assert len(it) == 1
        with api.Uahn(wfile, variant, precision="fp32", max_batch=1) as net:
            # the shim numbers its forwards 0, 1, ...: main model = pair index 0, iteration 1 = pair index 1 (seed 9)
            m, _, _ = net.infer_batch(frames[0:1], frames[1:2], pri[1:2].astype(np.float32), seed=9, first_pair=1)
        assert np.allclose(np.array(it[0][2:10], float), m[0], atol=1e-6), variant
    # a file without a record falls back to the 2-block schedule (and says nothing alarming)
    out = subprocess.run([exe, wfile, str(tmp_path / "frames.u8"), "2", str(tmp_path / "priors.f64"), "0", wfile],
                         check=True, capture_output=True, text=True)
    it = [l.split() for l in out.stdout.splitlines() if l.startswith("ITER")]
    with api.Uahn(wfile, "prior2", precision="fp32", max_batch=1) as net:
        m, _, _ = net.infer_batch(frames[0:1], frames[1:2], pri[1:2].astype(np.float32), seed=9, first_pair=1)
    assert np.allclose(np.array(it[0][2:10], float), m[0], atol=1e-6)

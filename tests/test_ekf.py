"""SURVEY §8f row 1 — the EKF measurement update behind the C ABI (host only) vs the numpy restatement and analytic
known answers.  CPU tests; the IEKF loop around uahn_infer needs a GPU."""
import numpy as np
import pytest

from cuahn_vio_b200 import synthetic as S
from oracle import ekf_oracle as E


@pytest.fixture(scope="module")
def api():
    from cuahn_vio_b200 import build
    build.build()
    from cuahn_vio_b200 import api as a
    a.load_library()
    return a


def _random_state(rng, scale=1e-2):
    A = rng.standard_normal((27, 27))
    P = scale * (A @ A.T) / 27 + 1e-4 * np.eye(27)
    imu = np.concatenate([rng.standard_normal(3), [0, 0, 0, 0], rng.standard_normal(3), 0.1 * rng.standard_normal(6)])
    q = rng.standard_normal(4)
    imu[3:7] = q / np.linalg.norm(q)
    offsets = 0.05 * rng.standard_normal((4, 3))
    return imu, offsets, P


def _random_meas(rng):
    mean = 8.0 * rng.standard_normal(8)
    cov = np.zeros((8, 8))
    for c in range(4):                           # block-diagonal 2x2 like combined_stu_model (model_to_trace.py:313-317)
        B = rng.standard_normal((2, 2))
        cov[2 * c:2 * c + 2, 2 * c:2 * c + 2] = B @ B.T + 0.05 * np.eye(2)
    return mean, cov


@pytest.mark.parametrize("update_offset", [True, False])
def test_update_matches_restatement(api, update_offset):
    rng = np.random.default_rng(5)
    for trial in range(20):
        imu, off, P = _random_state(rng)
        mean, cov = _random_meas(rng)
        st = api.EkfState.from_arrays(imu, off, P)
        _, prop = api.ekf_prior_px(st)
        assert np.array_equal(prop, off[:, :2].reshape(8))
        api.ekf_update(st, mean, cov, prop, update_offset, 10.0)
        imu2, off2, P2 = st.arrays()
        ri, ro, rP = E.update(imu, off, P, mean, cov, prop, update_offset, 10.0)
        assert np.abs(imu2 - ri).max() < 1e-12 and np.abs(off2 - ro).max() < 1e-12
        assert np.abs(P2 - rP).max() < 1e-12 * max(1.0, np.abs(rP).max())
        if not update_offset:
            assert np.array_equal(off2, off)


def test_known_answers(api):
    imu = np.zeros(16); imu[3] = 1.0
    off = np.zeros((4, 3))
    # (1) offsets-only prior covariance, near-perfect measurement: the offsets jump onto the measurement / 159.5
    P = np.zeros((27, 27))
    for i in (15, 16, 18, 19, 21, 22, 24, 25):
        P[i, i] = 1.0
    mean = np.arange(1, 9, dtype=np.float64)
    st = api.EkfState.from_arrays(imu, off, P)
    api.ekf_update(st, mean, 1e-12 * np.eye(8), np.zeros(8), True, 10.0)
    _, off2, P2 = st.arrays()
    assert np.abs(off2[:, :2].reshape(8) - mean / 159.5).max() < 1e-12 and np.abs(off2[:, 2]).max() == 0
    assert np.abs(P2).max() < 1e-12
    # (2) hopeless measurement (huge covariance): nothing moves
    st = api.EkfState.from_arrays(imu, off, P)
    api.ekf_update(st, mean, 1e18 * np.eye(8), np.zeros(8), True, 10.0)
    imu3, off3, P3 = st.arrays()
    assert np.abs(off3).max() < 1e-9 and np.abs(P3 - P).max() < 1e-9 and np.abs(imu3 - imu).max() < 1e-12
    # (3) equal prior and measurement variance: the offset moves half way; units: K_net_Cov * cov / 159.5^2
    st = api.EkfState.from_arrays(imu, off, P)
    api.ekf_update(st, mean, (159.5 ** 2 / 10.0) * np.eye(8), np.zeros(8), True, 10.0)
    _, off4, P4 = st.arrays()
    assert np.abs(off4[:, :2].reshape(8) - 0.5 * mean / 159.5).max() < 1e-12
    assert abs(P4[15, 15] - 0.5) < 1e-12
    # (4) reset: offsets zero, only the IMU block of the covariance survives
    rng = np.random.default_rng(1)
    imu, off, P = _random_state(rng)
    st = api.EkfState.from_arrays(imu, off, P)
    api.load_library().uahn_ekf_reset_offsets(st)
    _, off5, P5 = st.arrays()
    ro, rP = E.reset_4pt_offset(off, P)
    assert np.array_equal(off5, ro) and np.array_equal(P5, rP)
    # (5) a pure rotation increment turns the attitude by that angle, unit norm kept
    P = np.zeros((27, 27)); P[15, 15] = 1.0; P[3, 15] = P[15, 3] = 0.2; P[3, 3] = 1.0     # theta_x correlated with f_ul.u
    st = api.EkfState.from_arrays(np.concatenate([np.zeros(3), [1, 0, 0, 0], np.zeros(9)]), np.zeros((4, 3)), P)
    m = np.zeros(8); m[0] = 159.5 * 0.3
    api.ekf_update(st, m, 1e-12 * np.eye(8), np.zeros(8), False, 10.0)
    imu6, _, _ = st.arrays()
    ang = 0.2 * 0.3                                # K[3] = P[3,15] / P[15,15]
    assert np.abs(imu6[3:7] - np.array([np.cos(ang / 2), np.sin(ang / 2), 0, 0])).max() < 1e-12
    # singular innovation covariance is reported, not inverted
    st = api.EkfState.from_arrays(imu, off, np.zeros((27, 27)))
    with pytest.raises(api.UahnError):
        api.ekf_update(st, m, np.zeros((8, 8)), np.zeros(8), True, 10.0)


@pytest.mark.gpu
def test_iekf_frame_loop():
    """VioManager.cpp:227-275 on the GPU path: prior from the state, network call, gated update, offset reset."""
    from cuahn_vio_b200 import api, weights
    wfile = weights.synthetic_weights_file(0)
    frames, gt, _ = S.synthetic_sequence(4, seed=5)
    rng = np.random.default_rng(2)
    imu, _, P = _random_state(rng, scale=1e-4)
    off = np.zeros((4, 3)); off[:, :2] = gt[0] / 159.5            # the propagated offsets = a perfect prior
    with api.Uahn(wfile, "prior3", precision="fp32", max_batch=1) as net:
        net.load_image(frames[0], 0.0)
        net.load_image(frames[1], 1.0)
        # gate closed (image count <= min_images): the network runs, the state only has its offsets reset
        st = api.EkfState.from_arrays(imu, off, P)
        mean, cov = net.iekf_frame(st, max_iter=1, min_images=10, seed=3, pair_index=0)
        imu1, off1, P1 = st.arrays()
        ro, rP = E.reset_4pt_offset(off, P)
        assert np.array_equal(imu1, imu) and np.array_equal(off1, ro) and np.array_equal(P1, rP)
        m_ref, c_ref, _ = net.infer(off[:, :2].reshape(8) * 159.5, seed=3, pair_index=0)
        assert np.array_equal(mean, m_ref) and np.array_equal(cov, c_ref)
        # gate open, 2 IEKF iterations: iteration 0 updates offsets too, iteration 1 re-linearises around them
        st = api.EkfState.from_arrays(imu, off, P)
        net.iekf_frame(st, max_iter=2, min_images=0, seed=3, pair_index=0)
        imu2, off2, P2 = st.arrays()
        i_r, o_r, P_r = E.update(imu, off, P, m_ref, c_ref, off[:, :2].reshape(8), True, 10.0)
        m2, c2, _ = net.infer(o_r[:, :2].reshape(8) * 159.5, seed=3, pair_index=1)   # iteration `it` draws pair index first_pair_index + it
        i_r, o_r, P_r = E.update(i_r, o_r, P_r, m2, c2, o_r[:, :2].reshape(8), False, 10.0)
        o_r, P_r = E.reset_4pt_offset(o_r, P_r)
        assert np.abs(imu2 - i_r).max() < 1e-10 and np.array_equal(off2, o_r)
        assert np.abs(P2 - P_r).max() < 1e-12

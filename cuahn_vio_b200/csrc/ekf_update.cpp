// Host-side consumer of the UAHN outputs: the EKF measurement update of CUAHN-VIO and the IEKF loop
// around the network call, Eigen-free.  Reference: cuahn_ros/cuahn/src/update/UpdaterHNet.cpp:28-61,
// cuahn_ros/cuahn/src/update/UpdaterHNet.h:56-66, cuahn_ros/cuahn/src/state/State.cpp:101-111,
// cuahn_ros/cuahn/src/core/VioManager.cpp:227-275, ov_core/src/utils/quat_ops.h:141-145,479-484,526-538.
#include <cmath>
#include <cstring>

#include "../../include/uahn_ekf.h"

namespace {

constexpr int D = UAHN_EKF_DIM;
// rows of the measurement Jacobian H (UpdaterHNet.h:58-62): the (u, v) components of the four offsets
constexpr int MEAS[8] = {15, 16, 18, 19, 21, 22, 24, 25};

// In-place inverse of an 8x8 matrix by Gauss-Jordan elimination with partial pivoting.
bool invert8(double* a) {
  double inv[64];
  for (int i = 0; i < 64; ++i) inv[i] = (i / 8 == i % 8) ? 1.0 : 0.0;
  for (int c = 0; c < 8; ++c) {
    int piv = c;
    for (int r = c + 1; r < 8; ++r)
      if (std::fabs(a[r * 8 + c]) > std::fabs(a[piv * 8 + c])) piv = r;
    if (!(std::fabs(a[piv * 8 + c]) > 0.0)) return false;   // singular (or NaN)
    if (piv != c)
      for (int k = 0; k < 8; ++k) {
        const double t = a[c * 8 + k]; a[c * 8 + k] = a[piv * 8 + k]; a[piv * 8 + k] = t;
        const double u = inv[c * 8 + k]; inv[c * 8 + k] = inv[piv * 8 + k]; inv[piv * 8 + k] = u;
      }
    const double d = 1.0 / a[c * 8 + c];
    for (int k = 0; k < 8; ++k) { a[c * 8 + k] *= d; inv[c * 8 + k] *= d; }
    for (int r = 0; r < 8; ++r) {
      if (r == c) continue;
      const double f = a[r * 8 + c];
      if (f == 0.0) continue;
      for (int k = 0; k < 8; ++k) { a[r * 8 + k] -= f * a[c * 8 + k]; inv[r * 8 + k] -= f * inv[c * 8 + k]; }
    }
  }
  std::memcpy(a, inv, sizeof(inv));
  return true;
}

}  // namespace

extern "C" {

int uahn_ekf_prior_px(const uahn_ekf_state* s, double* prior_px8, double* propagated8) {
  if (!s || !prior_px8) return UAHN_ERR_INVALID;
  for (int c = 0; c < 4; ++c)
    for (int k = 0; k < 2; ++k) {
      const double v = s->offset[c][k];                        // VioManager.cpp:230-233
      if (propagated8) propagated8[2 * c + k] = v;
      prior_px8[2 * c + k] = v * UAHN_FOCAL_PX;                // :234
    }
  return UAHN_OK;
}

int uahn_ekf_update(uahn_ekf_state* s, const double* mean_px8, const double* cov_px64, const double* propagated8,
                    int update_offset, double K_net_Cov) {
  if (!s || !mean_px8 || !cov_px64 || !propagated8) return UAHN_ERR_INVALID;
  double* P = s->cov;
  // S = H P H^T + Hn (K_net_Cov * Cov / 25440.25) Hn^T, Hn = I (UpdaterHNet.cpp:31, UpdaterHNet.h:64)
  double Sinv[64];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) Sinv[i * 8 + j] = P[MEAS[i] * D + MEAS[j]] + K_net_Cov * cov_px64[i * 8 + j] / 25440.25;
  if (!invert8(Sinv)) return UAHN_ERR_INVALID;
  // K = P H^T S^-1  (27 x 8)
  double K[D * 8];
  for (int r = 0; r < D; ++r)
    for (int j = 0; j < 8; ++j) {
      double acc = 0.0;
      for (int i = 0; i < 8; ++i) acc += P[r * D + MEAS[i]] * Sinv[i * 8 + j];
      K[r * 8 + j] = acc;
    }
  double inno[8];
  for (int i = 0; i < 8; ++i) inno[i] = mean_px8[i] / UAHN_FOCAL_PX - propagated8[i];            // :33
  // P <- (I - K H) P = P - K (H P)   (:36)
  double HP[8 * D];
  for (int i = 0; i < 8; ++i) std::memcpy(HP + i * D, P + MEAS[i] * D, D * sizeof(double));
  for (int r = 0; r < D; ++r)
    for (int c = 0; c < D; ++c) {
      double acc = 0.0;
      for (int i = 0; i < 8; ++i) acc += K[r * 8 + i] * HP[i * D + c];
      P[r * D + c] -= acc;
    }
  // state increment (:39-44); without update_offset only the 15 IMU rows are used
  double dx[D];
  const int rows = update_offset ? D : 15;
  for (int r = 0; r < rows; ++r) {
    double acc = 0.0;
    for (int i = 0; i < 8; ++i) acc += K[r * 8 + i] * inno[i];
    dx[r] = acc;
  }
  double* x = s->imu;
  for (int k = 0; k < 3; ++k) x[k] += dx[k];                                                       // :47
  {   // :48  q <- quatnorm(Ham_quat_update(dtheta) * q)
    const double* w = dx + 3;
    const double angle = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    // the reference divides by the angle unconditionally (quat_ops.h:530) and would produce NaN for an exactly
    // zero increment; the limit sin(a/2)/a -> 1/2 is used here instead
    const double k = angle > 0.0 ? std::sin(angle * 0.5) / angle : 0.5;
    const double d0 = k * w[0], d1 = k * w[1], d2 = k * w[2], c = std::cos(angle * 0.5);
    // qR = [[c, -d^T], [d, c I + skew_x(-d)]]  (quat_ops.h:532-535), skew_x(v) = [0 -vz vy; vz 0 -vx; -vy vx 0]
    const double M[16] = {c, -d0, -d1, -d2,
                          d0, c, d2, -d1,
                          d1, -d2, c, d0,
                          d2, d1, -d0, c};
    const double* q = x + 3;
    double qn[4];
    for (int r = 0; r < 4; ++r) qn[r] = M[r * 4] * q[0] + M[r * 4 + 1] * q[1] + M[r * 4 + 2] * q[2] + M[r * 4 + 3] * q[3];
    if (qn[3] < 0) for (int r = 0; r < 4; ++r) qn[r] = -qn[r];                                    // quat_ops.h:480-482 (sic: index 3)
    const double n = std::sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
    for (int r = 0; r < 4; ++r) x[3 + r] = qn[r] / n;
  }
  for (int k = 0; k < 3; ++k) x[7 + k] += dx[6 + k];                                               // :49
  for (int k = 0; k < 3; ++k) x[10 + k] += dx[9 + k];                                              // :50
  for (int k = 0; k < 3; ++k) x[13 + k] += dx[12 + k];                                             // :51
  if (update_offset)                                                                               // :55-60
    for (int c = 0; c < 4; ++c)
      for (int k = 0; k < 3; ++k) s->offset[c][k] += dx[15 + 3 * c + k];
  return UAHN_OK;
}

int uahn_ekf_reset_offsets(uahn_ekf_state* s) {
  if (!s) return UAHN_ERR_INVALID;
  std::memset(s->offset, 0, sizeof(s->offset));
  for (int r = 0; r < D; ++r)
    for (int c = 0; c < D; ++c)
      if (r >= 15 || c >= 15) s->cov[r * D + c] = 0.0;
  return UAHN_OK;
}

int uahn_ekf_iekf_frame(uahn_handle* h, uahn_handle* h_iter, uahn_ekf_state* s, int max_iter, double K_net_Cov,
                        int min_images, int use_measurement, const uahn_rng* rng, double* mean_px8, double* cov_px64) {
  if (!h || !s || max_iter < 1) return UAHN_ERR_INVALID;
  double mean[8] = {0}, cov[64] = {0};
  for (int it = 0; it < max_iter; ++it) {                                                          // VioManager.cpp:227
    double prior_px[8], propagated[8];
    uahn_ekf_prior_px(s, prior_px, propagated);                                                    // :230-234
    uahn_handle* net = (it > 0 && h_iter) ? h_iter : h;                                            // HomographyNet.cpp:209
    // fresh MC-dropout masks for every forward (model_to_trace.py:266-273): iteration `it` of this frame uses pair index
    // first_pair_index + it (the caller advances first_pair_index by max_iter per frame); a NULL rng lets each handle
    // number its own calls
    uahn_rng it_rng{};
    if (rng) { it_rng = *rng; it_rng.first_pair_index += (uint64_t)it; }
    int rc = uahn_infer(net, prior_px, rng ? &it_rng : nullptr, mean, cov, nullptr);               // :236
    if (rc == UAHN_ERR_STATE) continue;        // "Only has one image": outputs untouched, loop goes on (HomographyNet.cpp:155-158)
    if (rc) return rc;
    if (use_measurement && uahn_image_count(h) > min_images) {                                     // :257
      rc = uahn_ekf_update(s, mean, cov, propagated, it != max_iter - 1, K_net_Cov);               // :260-264
      if (rc) return rc;
    }
  }
  uahn_ekf_reset_offsets(s);                                                                       // :275
  if (mean_px8) std::memcpy(mean_px8, mean, sizeof(mean));
  if (cov_px64) std::memcpy(cov_px64, cov, sizeof(cov));
  return UAHN_OK;
}

}  // extern "C"

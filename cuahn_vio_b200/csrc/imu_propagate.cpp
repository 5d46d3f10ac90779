// IMU propagation of the CUAHN-VIO filter state, Eigen-free fp64 host code.
// Reference: cuahn_ros/cuahn/src/state/Propagator.cpp:28-79 (propagate_with_imu), :80-180 (select_imu_readings),
// :183-339 (predict_and_compute), :342-363 (predict_mean_discrete); Propagator.h:50-103,179-195;
// cuahn_ros/cuahn/src/state/StateHelper.cpp:28-32; State.h:110-113; ov_core/src/utils/quat_ops.h:141-145,479-484,
// 526-550,573-588.  Products are evaluated left to right like the Eigen expressions they restate.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/uahn_ekf.h"

namespace {

constexpr int D = UAHN_EKF_DIM;
struct V3 { double v[3]; };
struct M3 { double m[9]; };

inline M3 skew(const V3& w) { return {{0, -w.v[2], w.v[1], w.v[2], 0, -w.v[0], -w.v[1], w.v[0], 0}}; }   // quat_ops.h:141-145
inline M3 eye() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline M3 mul(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0;
      for (int k = 0; k < 3; ++k) acc += a.m[i * 3 + k] * b.m[k * 3 + j];
      c.m[i * 3 + j] = acc;
    }
  return c;
}
inline V3 mul(const M3& a, const V3& b) {
  V3 c;
  for (int i = 0; i < 3; ++i) c.v[i] = a.m[i * 3] * b.v[0] + a.m[i * 3 + 1] * b.v[1] + a.m[i * 3 + 2] * b.v[2];
  return c;
}
inline M3 tr(const M3& a) { return {{a.m[0], a.m[3], a.m[6], a.m[1], a.m[4], a.m[7], a.m[2], a.m[5], a.m[8]}}; }
inline M3 add(const M3& a, const M3& b) { M3 c; for (int i = 0; i < 9; ++i) c.m[i] = a.m[i] + b.m[i]; return c; }
inline M3 sub(const M3& a, const M3& b) { M3 c; for (int i = 0; i < 9; ++i) c.m[i] = a.m[i] - b.m[i]; return c; }
inline M3 scale(double s, const M3& a) { M3 c; for (int i = 0; i < 9; ++i) c.m[i] = s * a.m[i]; return c; }
inline V3 add(const V3& a, const V3& b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]}}; }
inline V3 sub(const V3& a, const V3& b) { return {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; }
inline V3 scale(double s, const V3& a) { return {{s * a.v[0], s * a.v[1], s * a.v[2]}}; }
inline double dot(const V3& a, const V3& b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
inline M3 outer(const V3& a, const V3& b) {
  M3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.m[i * 3 + j] = a.v[i] * b.v[j];
  return c;
}
inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }

// quat_ops.h:546-550 (Hamilton, scalar first): R = (q0^2 - |qv|^2) I + 2 qv qv^T + 2 q0 [qv]x
inline M3 ham_quat_2_rot(const double* q) {
  const V3 qv{{q[1], q[2], q[3]}};
  return add(add(scale(q[0] * q[0] - (q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), eye()), scale(2.0, outer(qv, qv))),
             scale(2.0 * q[0], skew(qv)));
}
// quat_ops.h:582-588
inline void rotvec_2_ham_quat(const V3& rv, double* q) {
  const double n = norm(rv);
  const double k = n > 0.0 ? std::sin(n * 0.5) / n : 0.5;      // (the reference divides by the norm unconditionally)
  q[0] = std::cos(n * 0.5);
  for (int i = 0; i < 3; ++i) q[1 + i] = k * rv.v[i];
}
// quat_ops.h:573-580
inline M3 jr_theta(const V3& th) {
  const double n = norm(th);
  if (!(n > 0.0)) return eye();                                 // limit of the reference's expression
  const M3 S = skew(th);
  return add(sub(eye(), scale((1 - std::cos(n)) / (n * n), S)), scale((n - std::sin(n)) / (n * n * n), mul(S, S)));
}
// quat_ops.h:526-538 then :479-484
inline void quat_update(const V3& rot_vec, const double* q, double* out) {
  const double angle = norm(rot_vec);
  const double k = angle > 0.0 ? std::sin(angle * 0.5) / angle : 0.5;
  const double d0 = k * rot_vec.v[0], d1 = k * rot_vec.v[1], d2 = k * rot_vec.v[2], c = std::cos(angle * 0.5);
  const double M[16] = {c, -d0, -d1, -d2, d0, c, d2, -d1, d1, -d2, c, d0, d2, d1, -d0, c};
  double qn[4];
  for (int r = 0; r < 4; ++r) qn[r] = M[r * 4] * q[0] + M[r * 4 + 1] * q[1] + M[r * 4 + 2] * q[2] + M[r * 4 + 3] * q[3];
  if (qn[3] < 0) for (int r = 0; r < 4; ++r) qn[r] = -qn[r];
  const double n = std::sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
  for (int r = 0; r < 4; ++r) out[r] = qn[r] / n;
}

inline void put(double* F, int ld, int r0, int c0, const M3& b) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) F[(r0 + i) * ld + c0 + j] = b.m[i * 3 + j];
}
inline M3 get(const double* F, int ld, int r0, int c0) {
  M3 b;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) b.m[i * 3 + j] = F[(r0 + i) * ld + c0 + j];
  return b;
}

uahn_imu_sample interpolate(const uahn_imu_sample& a, const uahn_imu_sample& b, double t) {   // Propagator.h:179-189
  const double lambda = (t - a.t) / (b.t - a.t);
  uahn_imu_sample d;
  d.t = t;
  for (int i = 0; i < 3; ++i) {
    d.am[i] = (1 - lambda) * a.am[i] + lambda * b.am[i];
    d.wm[i] = (1 - lambda) * a.wm[i] + lambda * b.wm[i];
  }
  return d;
}

std::vector<uahn_imu_sample> select(const uahn_imu_sample* imu, int n, double time0, double time1) {   // Propagator.cpp:80-180
  std::vector<uahn_imu_sample> prop;
  if (n <= 0) return prop;
  for (int i = 0; i < n - 1; ++i) {
    if (imu[i + 1].t > time0 && imu[i].t < time0) {            // start of the integration period: split
      prop.push_back(interpolate(imu[i], imu[i + 1], time0));
      continue;
    }
    if (imu[i].t >= time0 && imu[i + 1].t <= time1) {          // middle
      prop.push_back(imu[i]);
      continue;
    }
    if (imu[i + 1].t > time1) {                                // end: split the next measurement at time1
      if (imu[i].t > time1 && i == 0) {
        break;
      } else if (imu[i].t > time1) {
        prop.push_back(interpolate(imu[i - 1], imu[i], time1));
      } else {
        prop.push_back(imu[i]);
      }
      if (prop.back().t != time1) prop.push_back(interpolate(imu[i], imu[i + 1], time1));
      break;
    }
  }
  if (prop.empty()) return prop;
  for (size_t i = 0; i + 1 < prop.size(); ++i)                  // no zero-dt intervals
    if (std::fabs(prop[i + 1].t - prop[i].t) < 1e-12) {
      prop.erase(prop.begin() + i);
      --i;
    }
  return prop;
}

struct Cfg {
  M3 cRi;
  V3 it;
  double q_diag[15];
  V3 gravity;
  bool imu_avg;
};

Cfg resolve(const uahn_propagator_config* c) {
  Cfg r;
  std::memcpy(r.cRi.m, c->c_R_i, sizeof(r.cRi.m));
  std::memcpy(r.it.v, c->i_t_i2c, sizeof(r.it.v));
  const double sw = c->sigma_w > 0 ? c->sigma_w : 1.6968e-04, sa = c->sigma_a > 0 ? c->sigma_a : 2.0000e-3;   // Propagator.h:50-71
  const double swb = c->sigma_wb > 0 ? c->sigma_wb : 1.9393e-05, sab = c->sigma_ab > 0 ? c->sigma_ab : 3.0000e-03;
  for (int i = 0; i < 3; ++i) {                                                                                // Propagator.h:93-97
    r.q_diag[i] = std::pow(sw, 2);
    r.q_diag[3 + i] = std::pow(sa, 2);
    r.q_diag[6 + i] = std::pow(sab, 2);
    r.q_diag[9 + i] = std::pow(swb, 2);
    r.q_diag[12 + i] = 1.0e-04;
  }
  r.gravity = {{0.0, 0.0, -(c->gravity_mag > 0 ? c->gravity_mag : 9.81)}};                                     // Propagator.h:100
  r.imu_avg = c->imu_avg != 0;
  return r;
}

// Propagator.cpp:183-363
void predict_and_compute(const Cfg& c, uahn_ekf_state* s, const uahn_imu_sample& dm, const uahn_imu_sample& dp, double* F,
                         double* Fw) {
  const double dt = dp.t - dm.t;
  const V3 pos{{s->imu[0], s->imu[1], s->imu[2]}}, vel{{s->imu[7], s->imu[8], s->imu[9]}};
  const V3 ba{{s->imu[10], s->imu[11], s->imu[12]}}, bg{{s->imu[13], s->imu[14], s->imu[15]}};
  const V3 w1 = sub({{dm.wm[0], dm.wm[1], dm.wm[2]}}, bg), a1 = sub({{dm.am[0], dm.am[1], dm.am[2]}}, ba);
  const V3 w2 = sub({{dp.wm[0], dp.wm[1], dp.wm[2]}}, bg), a2 = sub({{dp.am[0], dp.am[1], dp.am[2]}}, ba);
  const V3 w_hat = c.imu_avg ? scale(.5, add(w1, w2)) : w2, a_hat = c.imu_avg ? scale(.5, add(a1, a2)) : a2;   // :198-204
  const M3 Rot = ham_quat_2_rot(s->imu + 3), RotT = tr(Rot);
  const V3 muw{{0.0, 0.0, -1.0}}, ez{{0.0, 0.0, 1.0}};
  const M3 I = eye();
  // :213-216
  const V3 wc = mul(c.cRi, w_hat);
  const V3 vc = mul(c.cRi, add(vel, mul(skew(w_hat), c.it)));
  const V3 muc = mul(mul(c.cRi, RotT), muw);
  const double dc = mul(Rot, add(pos, c.it)).v[2];
  static const double CAM[4][3] = {{-1.0, -0.69906, 1.0}, {-1.0, 0.69906, 1.0}, {1.0, 0.69906, 1.0}, {1.0, -0.69906, 1.0}};   // State.h:110-113
  V3 pt[4];
  for (int k = 0; k < 4; ++k) pt[k] = {{CAM[k][0] + s->offset[k][0], CAM[k][1] + s->offset[k][1], CAM[k][2] + s->offset[k][2]}};

  // ---- predict_mean_discrete (:342-363)
  double new_q[4];
  quat_update(scale(dt, w_hat), s->imu + 3, new_q);
  const V3 new_v = add(vel, scale(dt, add(add(scale(-1.0, mul(skew(w_hat), vel)), a_hat), mul(RotT, c.gravity))));
  const V3 new_p = add(pos, scale(dt, add(scale(-1.0, mul(skew(w_hat), pos)), vel)));
  const M3 H = add(skew(wc), scale(1.0 / dc, outer(vc, muc)));
  V3 new_off[4];
  for (int k = 0; k < 4; ++k) {
    const M3 common = sub(I, outer(pt[k], ez));
    const V3 cur{{s->offset[k][0], s->offset[k][1], s->offset[k][2]}};
    new_off[k] = add(cur, scale(dt, mul(scale(-1.0, common), mul(H, pt[k]))));
  }

  // ---- Jacobians (:224-325)
  std::memset(F, 0, sizeof(double) * D * D);
  std::memset(Fw, 0, sizeof(double) * D * 15);
  const int P = 0, Q = 3, V = 6, BA = 9, BG = 12;
  put(F, D, P, P, sub(I, scale(dt, skew(w_hat))));
  put(F, D, P, V, scale(dt, I));
  put(F, D, P, BG, scale(-dt, skew(pos)));
  double dq[4];
  rotvec_2_ham_quat(scale(dt, w_hat), dq);
  put(F, D, Q, Q, tr(ham_quat_2_rot(dq)));
  put(F, D, Q, BG, scale(-dt, jr_theta(scale(dt, w_hat))));
  put(F, D, V, Q, scale(dt, skew(mul(RotT, c.gravity))));
  put(F, D, V, V, sub(I, scale(dt, skew(w_hat))));
  put(F, D, V, BA, scale(-dt, I));
  put(F, D, V, BG, scale(-dt, skew(vel)));
  put(F, D, BA, BA, I);
  put(F, D, BG, BG, I);

  const double scalar = dot(ez, vc) / dc;                                                      // :240-241
  const M3 J_f_df = scale(-dt, I);                                                             // :288
  const V3 J_dc_p = mul(tr(Rot), ez);                        // (ezT * Rot)^T                    :289
  const V3 J_dc_q = mul(tr(scale(-1.0, mul(Rot, skew(add(pos, c.it))))), ez);                  // :290
  const M3 J_muc_q = mul(c.cRi, skew(mul(RotT, muw)));                                         // :291
  const M3 J_vc_v = c.cRi, J_vc_bw = mul(c.cRi, skew(c.it)), J_wc_bw = scale(-1.0, c.cRi);     // Propagator.h:192-194
  const M3 Swc = skew(wc);
  const V3 ezSwc = mul(tr(Swc), ez);                         // (ezT * skew_x(wc))^T
  for (int k = 0; k < 4; ++k) {
    const V3& p = pt[k];
    const double ez_swc_p = dot(ezSwc, p), muc_p = dot(muc, p);
    // J_df_pt = [wc]x + vc muc^T / dc - (ezT [wc]x pt) I - pt (ezT [wc]x) - scalar ((muc^T pt) I + pt muc^T)   (:244-246)
    M3 J_df_pt = add(Swc, scale(1.0 / dc, outer(vc, muc)));
    J_df_pt = sub(J_df_pt, scale(ez_swc_p, I));
    J_df_pt = sub(J_df_pt, outer(p, ezSwc));
    J_df_pt = sub(J_df_pt, scale(scalar, add(scale(muc_p, I), outer(p, muc))));
    const M3 common = sub(I, outer(p, ez));                                                    // :247
    const V3 J_df_dc = mul(scale(1.0 / dc / dc * muc_p, scale(-1.0, common)), vc);             // :248
    const M3 J_df_vc = scale(1.0 / dc * muc_p, common);                                        // :249
    const M3 J_df_muc = scale(1.0 / dc, outer(mul(common, vc), p));                            // :250
    const M3 J_df_wc = scale(-1.0, mul(common, skew(p)));                                      // :251
    const int R = 15 + 3 * k;
    put(F, D, R, P, outer(mul(J_f_df, J_df_dc), J_dc_p));                                      // :294
    put(F, D, R, Q, mul(J_f_df, add(outer(J_df_dc, J_dc_q), mul(J_df_muc, J_muc_q))));         // :295
    put(F, D, R, V, mul(mul(J_f_df, J_df_vc), J_vc_v));                                        // :296
    put(F, D, R, BG, mul(J_f_df, add(mul(J_df_vc, J_vc_bw), mul(J_df_wc, J_wc_bw))));          // :297
    put(F, D, R, R, add(I, mul(J_f_df, J_df_pt)));                                             // :298
  }
  // Fw (:319-332)
  const M3 dtI = get(F, D, P, V);
  put(Fw, 15, P, 0, scale(-1.0, get(F, D, P, BG)));
  put(Fw, 15, P, 12, dtI);
  put(Fw, 15, Q, 0, scale(-1.0, get(F, D, Q, BG)));
  put(Fw, 15, V, 0, scale(-1.0, get(F, D, V, BG)));
  put(Fw, 15, V, 3, dtI);
  put(Fw, 15, BA, 6, dtI);
  put(Fw, 15, BG, 9, dtI);
  for (int k = 0; k < 4; ++k) put(Fw, 15, 15 + 3 * k, 0, scale(-1.0, get(F, D, 15 + 3 * k, BG)));

  // new mean (:335-345)
  for (int i = 0; i < 3; ++i) { s->imu[i] = new_p.v[i]; s->imu[7 + i] = new_v.v[i]; }
  for (int i = 0; i < 4; ++i) s->imu[3 + i] = new_q[i];
  for (int k = 0; k < 4; ++k)
    for (int i = 0; i < 3; ++i) s->offset[k][i] = new_off[k].v[i];
}

// StateHelper.cpp:28-32: P <- F P F^T + Fw Q Fw^T (Q diagonal)
void propagate_cov(const Cfg& c, uahn_ekf_state* s, const double* F, const double* Fw) {
  std::vector<double> FP(D * D), Pn(D * D);
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      double acc = 0;
      for (int k = 0; k < D; ++k) acc += F[i * D + k] * s->cov[k * D + j];
      FP[i * D + j] = acc;
    }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      double acc = 0;
      for (int k = 0; k < D; ++k) acc += FP[i * D + k] * F[j * D + k];
      double nq = 0;
      for (int k = 0; k < 15; ++k) nq += Fw[i * 15 + k] * c.q_diag[k] * Fw[j * 15 + k];
      Pn[i * D + j] = acc + nq;
    }
  std::memcpy(s->cov, Pn.data(), sizeof(double) * D * D);
}

}  // namespace

extern "C" {

int uahn_imu_select_readings(const uahn_imu_sample* imu, int n, double time0, double time1, uahn_imu_sample* out, int capacity,
                             int* n_out) {
  if (!imu || !out || !n_out || n < 0) return UAHN_ERR_INVALID;
  const std::vector<uahn_imu_sample> p = select(imu, n, time0, time1);
  if ((int)p.size() > capacity) return UAHN_ERR_INVALID;
  for (size_t i = 0; i < p.size(); ++i) out[i] = p[i];
  *n_out = (int)p.size();
  return UAHN_OK;
}

int uahn_imu_predict_and_compute(const uahn_propagator_config* cfg, uahn_ekf_state* s, const uahn_imu_sample* minus,
                                 const uahn_imu_sample* plus, double* F, double* Fw) {
  if (!cfg || !s || !minus || !plus || !F || !Fw) return UAHN_ERR_INVALID;
  predict_and_compute(resolve(cfg), s, *minus, *plus, F, Fw);
  return UAHN_OK;
}

int uahn_imu_propagate(const uahn_propagator_config* cfg, uahn_ekf_state* s, const uahn_imu_sample* imu, int n, double time0,
                       double time1, int* n_intervals) {
  if (!cfg || !s || !imu) return UAHN_ERR_INVALID;
  if (!(time1 > time0)) return UAHN_ERR_STATE;   // Propagator.cpp:32-42: same instant / backwards (the reference exits)
  const Cfg c = resolve(cfg);
  const std::vector<uahn_imu_sample> p = select(imu, n, time0, time1);
  std::vector<double> F(D * D), Fw(D * 15);
  int count = 0;
  if (p.size() > 1)                                                                            // :64-71
    for (size_t i = 0; i + 1 < p.size(); ++i, ++count) {
      predict_and_compute(c, s, p[i], p[i + 1], F.data(), Fw.data());
      propagate_cov(c, s, F.data(), Fw.data());
    }
  if (n_intervals) *n_intervals = count;
  return UAHN_OK;
}

}  // extern "C"

/*
 * uahn_ekf.h — host-side consumer of the UAHN outputs: the EKF measurement update and the IEKF loop glue of
 * CUAHN-VIO ("next" row 1).  Plain C ABI, Eigen-free, double precision, CPU only (27x27 algebra).
 *
 * Replaces, for a ROS-free replay harness or a binding that does not want Eigen:
 *   UpdaterHNet::update                  cuahn_ros/cuahn/src/update/UpdaterHNet.cpp:28-61  (H / Hn: UpdaterHNet.h:56-66)
 *   State::reset_4pt_offset              cuahn_ros/cuahn/src/state/State.cpp:101-111
 *   the IEKF loop of VioManager          cuahn_ros/cuahn/src/core/VioManager.cpp:227-275
 *
 * State layout = the reference's (State.cpp:31-91): error state [p(3) theta(3) v(3) ba(3) bg(3) f_ul(3) f_bl(3)
 * f_br(3) f_ur(3)] = 27; IMU value [p(3), q(4, Hamilton, as stored by IMU_CUAHN), v(3), ba(3), bg(3)] = 16;
 * the four 4-point offsets are 3-vectors in camera-normalised units (pixels / 159.5) of which the network measures
 * the first two components.
 */
#ifndef UAHN_EKF_H_
#define UAHN_EKF_H_

#include "uahn.h"

#ifdef __cplusplus
extern "C" {
#endif

#define UAHN_EKF_DIM 27
#define UAHN_FOCAL_PX 159.5 /* (320-1)/2 / tan(45 deg): CamBase.h:166-169, VioManager.cpp:234, UpdaterHNet.cpp:31 */

typedef struct uahn_ekf_state {
  double imu[16];                          /* p, q, v, ba, bg (IMU_CUAHN value order, UpdaterHNet.cpp:45-50)        */
  double offset[4][3];                     /* UL, BL, BR, UR (State.h: _offset_upperLeft ... _offset_upperRight)    */
  double cov[UAHN_EKF_DIM * UAHN_EKF_DIM]; /* State::_Cov, row-major                                                 */
} uahn_ekf_state;

/* prior_px[8] = the (u, v) components of the four offsets x 159.5 (VioManager.cpp:230-234);
 * propagated[8] (optional) receives the same values in normalised units (what update() takes as 4th argument). */
UAHN_API int uahn_ekf_prior_px(const uahn_ekf_state* s, double* prior_px8, double* propagated8);

/* UpdaterHNet::update: K = P H^T (H P H^T + K_net_Cov * Cov / 159.5^2)^-1, innovation = mean / 159.5 - propagated,
 * P <- (I - K H) P, state += K * innovation (quaternion: quatnorm(Ham_quat_update(dtheta) * q)).  update_offset = 0
 * leaves the four offsets untouched (they are about to be reset).  mean_px / cov_px are the network outputs in pixels /
 * pixels^2, cov_px row-major 8x8.  Returns UAHN_ERR_INVALID on a singular innovation covariance. */
UAHN_API int uahn_ekf_update(uahn_ekf_state* s, const double* mean_px8, const double* cov_px64, const double* propagated8,
                    int update_offset, double K_net_Cov);

/* State::reset_4pt_offset: offsets <- 0, covariance keeps only its IMU 15x15 block. */
UAHN_API int uahn_ekf_reset_offsets(uahn_ekf_state* s);

/* One camera frame of the IEKF loop (VioManager.cpp:227-275) on the pair currently held by the handle(s):
 * for it in [0, max_iter): prior <- state offsets x 159.5; uahn_infer on `h` (it == 0) or `h_iter` (it > 0: the
 * reference's second, "iterative" model, HomographyNet.cpp:209-230; NULL = reuse h); if (image count > min_images)
 * update with update_offset = (it != max_iter - 1); finally reset the offsets.  Both handles must have been given the
 * same frames with uahn_load_image.  mean_px8 / cov_px64 (optional) receive the last network output.  The
 * reference's timestamp gate (HNet->get_latest_inference_time() == time_stamp, :257) is the caller's: pass
 * use_measurement = 0 to skip the updates.  Needs a GPU (it calls uahn_infer). */
UAHN_API int uahn_ekf_iekf_frame(uahn_handle* h, uahn_handle* h_iter, uahn_ekf_state* s, int max_iter, double K_net_Cov,
                        int min_images, int use_measurement, const uahn_rng* rng, double* mean_px8, double* cov_px64);

/* ---- IMU propagation producing the prior ---------------------------------------------------
 * Replaces Propagator::select_imu_readings / interpolate_data (cuahn/src/state/Propagator.cpp:80-180, Propagator.h:179-189),
 * Propagator::predict_and_compute + predict_mean_discrete (Propagator.cpp:183-363) and StateHelper::propagate_Cov
 * (StateHelper.cpp:28-32): the planar-homography corner dynamics that turn IMU readings into the propagated 4-point
 * offsets (the network's prior) and their covariance. */
typedef struct uahn_imu_sample {
  double t;
  double wm[3]; /* gyroscope, rad/s  */
  double am[3]; /* accelerometer, m/s^2 */
} uahn_imu_sample;

typedef struct uahn_propagator_config {
  double c_R_i[9];    /* State::c_RotMtrx_i, row-major (rotation block of the camera extrinsics, State.cpp:93-96) */
  double i_t_i2c[3];  /* State::i_tVec_i2c                                                                           */
  double sigma_w, sigma_a, sigma_wb, sigma_ab; /* NoiseManager (Propagator.h:50-71); <= 0 selects the reference defaults */
  double gravity_mag; /* 9.81 (Propagator.h:100); <= 0 selects it                                                    */
  int imu_avg;        /* StateOptions::imu_avg (default true, StateOptions.h:39)                                     */
} uahn_propagator_config;

/* select_imu_readings: the readings that cover [time0, time1], first and last interpolated onto the interval ends.
 * Writes at most `capacity` samples to `out`, their count to *n_out (fewer than 2 = cannot propagate). */
UAHN_API int uahn_imu_select_readings(const uahn_imu_sample* imu, int n, double time0, double time1, uahn_imu_sample* out,
                             int capacity, int* n_out);

/* predict_and_compute for one IMU interval: advances the state mean (IMU value and the four offsets) and returns the
 * state-transition Jacobian F (27x27) and noise Jacobian Fw (27x15), row-major.  The covariance is NOT touched. */
UAHN_API int uahn_imu_predict_and_compute(const uahn_propagator_config* cfg, uahn_ekf_state* s, const uahn_imu_sample* minus,
                                 const uahn_imu_sample* plus, double* F, double* Fw);

/* propagate_with_imu (Propagator.cpp:28-79) between two (already time-offset-corrected) instants: select readings,
 * then for every interval predict_and_compute + P <- F P F^T + Fw Q Fw^T.  Returns the number of intervals integrated
 * in *n_intervals (optional). */
UAHN_API int uahn_imu_propagate(const uahn_propagator_config* cfg, uahn_ekf_state* s, const uahn_imu_sample* imu, int n,
                       double time0, double time1, int* n_intervals);

#ifdef __cplusplus
}
#endif
#endif /* UAHN_EKF_H_ */

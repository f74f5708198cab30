// SPDX-License-Identifier: Apache-2.0
//
// observers_core.cuh -- the spine's observer pipeline, per robot (scalar recurrences).
//
// Restates, in the order of spines/common/observers.h:23-44:
//   BaseOrientation   upkie/cpp/observers/BaseOrientation.h:29-148 (pitch from the IMU quaternion,
//                     base angular velocity from the IMU gyroscope)
//   FloorContact      upkie/cpp/observers/FloorContact.cpp:37-91 with
//   WheelContact      upkie/cpp/observers/WheelContact.cpp:19-48 (inertia regressor with hysteresis)
//   WheelOdometry     upkie/cpp/observers/WheelOdometry.cpp:16-64 (contact-gated, integrated)
// and the filter upkie/cpp/utils/low_pass_filter.h:21-38.
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/upkie_b200.h"

#if defined(__CUDACC__)
#define UPKIE_OBS_HD __host__ __device__ __forceinline__
#else
#define UPKIE_OBS_HD inline
#endif

namespace upkie_b200 {

UPKIE_OBS_HD float obs_sqrt(float x) { return sqrtf(x); }
UPKIE_OBS_HD double obs_sqrt(double x) { return sqrt(x); }
UPKIE_OBS_HD float obs_acos(float x) { return acosf(x); }
UPKIE_OBS_HD double obs_acos(double x) { return acos(x); }
UPKIE_OBS_HD float obs_abs(float x) { return fabsf(x); }
UPKIE_OBS_HD double obs_abs(double x) { return fabs(x); }

template <typename T>
struct ObserverParams {
  T dt;
  T cutoff_period, liftoff_inertia, min_touchdown_acceleration, min_touchdown_torque, touchdown_inertia;
  T upper_leg_torque_threshold;
  T signed_radius[2];  // left, right wheel
  T Rbi[9];            // rotation_base_to_imu
};

template <typename T>
struct WheelContactState {
  T velocity, abs_acceleration, abs_torque, inertia, contact;
};

template <typename T>
struct ObserverState {
  WheelContactState<T> wheel[2];
  T upper_leg_torque;
  T odom_position, odom_velocity;
};

template <typename T>
UPKIE_OBS_HD T obs_lpf(T prev_output, T cutoff_period, T new_input, T dt) {
  // low_pass_filter.h:35-37 (the cutoff_period > 2 dt guard is checked once, at create time)
  const T alpha = dt / cutoff_period;
  return prev_output + alpha * (new_input - prev_output);
}

// WheelContact::observe (WheelContact.cpp:19-48)
template <typename T>
UPKIE_OBS_HD void wheel_contact_observe(const ObserverParams<T>& P, WheelContactState<T>& w, T torque, T velocity) {
  const T prev_velocity = w.velocity;
  w.velocity = obs_lpf(w.velocity, P.cutoff_period, velocity, P.dt);
  const T new_acceleration = (w.velocity - prev_velocity) / P.dt;
  w.abs_acceleration = obs_lpf(w.abs_acceleration, P.cutoff_period, obs_abs(new_acceleration), P.dt);
  w.abs_torque = obs_lpf(w.abs_torque, P.cutoff_period, obs_abs(torque), P.dt);
  const bool in_contact = w.contact > T(0.5);
  if (!in_contact && (w.abs_acceleration < P.min_touchdown_acceleration || w.abs_torque < P.min_touchdown_torque)) return;
  w.inertia = w.abs_torque / (w.abs_acceleration + T(1e-4));
  if (w.inertia < P.liftoff_inertia) w.contact = T(0);
  else if (w.inertia > P.touchdown_inertia) w.contact = T(1);
}

// compute_pitch_frame_in_parent (BaseOrientation.h:73-92); R row-major
template <typename T>
UPKIE_OBS_HD T pitch_frame_in_parent(const T R[9]) {
  T sx = R[0], sy = R[3], sz = R[6];  // first column: sagittal axis
  const T n = obs_sqrt(sx * sx + sy * sy + sz * sz);
  sx /= n; sy /= n; sz /= n;
  T hx = sx, hy = sy;  // heading = sagittal - sagittal.z * e_z
  const T hn = obs_sqrt(hx * hx + hy * hy);
  hx /= hn; hy /= hn;
  if (R[8] < T(0)) { hx = -hx; hy = -hy; }
  const T sign = (sz < T(0)) ? T(1) : T(-1);
  T cos_pitch = sx * hx + sy * hy;
  if (cos_pitch < T(-1)) cos_pitch = T(-1);
  else if (cos_pitch > T(1)) cos_pitch = T(1);
  return sign * obs_acos(cos_pitch);
}

// quaternion (w, x, y, z) -> rotation matrix, as Eigen::Quaterniond::toRotationMatrix
template <typename T>
UPKIE_OBS_HD void obs_quat_to_rot(const T q[4], T R[9]) {
  const T w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

// One cycle of the observer pipeline. Inputs: the spine observation row of this robot (imu.orientation,
// imu.angular_velocity, servo.*.torque / velocity). Output row: UPKIE_OBSV_* layout.
template <typename T>
UPKIE_OBS_HD void observers_step(const ObserverParams<T>& P, ObserverState<T>& st, const T* spine, T* out) {
  // BaseOrientation: R_base_to_world = R_ars_to_world * R(q_imu_in_ars) * R_base_to_imu (BaseOrientation.h:29-37)
  T Ria[9];
  obs_quat_to_rot(spine + UPKIE_SP_IMU_QUAT, Ria);
  T Rwa_Ria[9];  // diag(1, -1, -1) * Ria
  for (int j = 0; j < 3; ++j) { Rwa_Ria[j] = Ria[j]; Rwa_Ria[3 + j] = -Ria[3 + j]; Rwa_Ria[6 + j] = -Ria[6 + j]; }
  T Rbw[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Rbw[3 * i + j] = Rwa_Ria[3 * i] * P.Rbi[j] + Rwa_Ria[3 * i + 1] * P.Rbi[3 + j] + Rwa_Ria[3 * i + 2] * P.Rbi[6 + j];
  out[UPKIE_OBSV_PITCH] = pitch_frame_in_parent(Rbw);
  // angular velocity base in base = R_base_to_imu^T * angular_velocity_imu_in_imu (BaseOrientation.h:144-148)
  const T* wi = spine + UPKIE_SP_IMU_ANGVEL;
  for (int j = 0; j < 3; ++j) out[UPKIE_OBSV_ANGVEL + j] = P.Rbi[j] * wi[0] + P.Rbi[3 + j] * wi[1] + P.Rbi[6 + j] * wi[2];
  for (int i = 0; i < 9; ++i) out[UPKIE_OBSV_ROT + i] = Rbw[i];

  // FloorContact::read (FloorContact.cpp:37-91)
  const int wheel_joint[2] = {2, 5};
  bool at_least_one = false;
  for (int k = 0; k < 2; ++k) {
    const T* servo = spine + UPKIE_SP_SERVO + wheel_joint[k] * UPKIE_OBS_KEYS;
    wheel_contact_observe(P, st.wheel[k], servo[UPKIE_OBS_TORQUE], servo[UPKIE_OBS_VELOCITY]);
    if (st.wheel[k].contact > T(0.5)) at_least_one = true;
  }
  T squared = 0;
  const int upper_leg[4] = {0, 1, 3, 4};
  for (int k = 0; k < 4; ++k) {
    const T t = spine[UPKIE_SP_SERVO + upper_leg[k] * UPKIE_OBS_KEYS + UPKIE_OBS_TORQUE];
    squared += t * t;
  }
  st.upper_leg_torque = obs_lpf(st.upper_leg_torque, T(0.01), obs_sqrt(squared), P.dt);  // kTorqueCutoffPeriod
  const bool contact = at_least_one || (st.upper_leg_torque > P.upper_leg_torque_threshold);
  out[UPKIE_OBSV_CONTACT] = contact ? T(1) : T(0);
  out[UPKIE_OBSV_WHEEL_CONTACT + 0] = st.wheel[0].contact;
  out[UPKIE_OBSV_WHEEL_CONTACT + 1] = st.wheel[1].contact;
  out[UPKIE_OBSV_LEG_TORQUE] = st.upper_leg_torque;
  out[UPKIE_OBSV_WHEEL_INERTIA + 0] = st.wheel[0].inertia;
  out[UPKIE_OBSV_WHEEL_INERTIA + 1] = st.wheel[1].inertia;

  // WheelOdometry::read (WheelOdometry.cpp:16-48): only while the floor contact holds
  if (contact) {
    T sum = 0;
    int nb = 0;
    for (int k = 0; k < 2; ++k) {
      if (st.wheel[k].contact > T(0.5)) {
        sum += P.signed_radius[k] * spine[UPKIE_SP_SERVO + wheel_joint[k] * UPKIE_OBS_KEYS + UPKIE_OBS_VELOCITY];
        ++nb;
      }
    }
    st.odom_velocity = nb > 0 ? sum / T(nb) : T(0);
    st.odom_position += st.odom_velocity * P.dt;
  }
  out[UPKIE_OBSV_ODOM_POS] = st.odom_position;
  out[UPKIE_OBSV_ODOM_VEL] = st.odom_velocity;
}

}  // namespace upkie_b200

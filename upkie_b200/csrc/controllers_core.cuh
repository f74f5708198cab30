// SPDX-License-Identifier: Apache-2.0
//
// controllers_core.cuh -- the spine's "wheel_balancer" controller pipeline for one robot
// (spines/common/controllers.h:24-44): WheelStopper then WheelBalancer, each read() then write()
// (upkie/cpp/controllers/ControllerPipeline.cpp:20-33). __host__ __device__ so that tests/hostsim can run
// the very same arithmetic on the CPU; the fp64 restatement the parity tests compare against is in oracle/.
#pragma once

#include "sim_core.cuh"

namespace upkie_b200 {

template <typename T>
struct WheelBalancerParams {
  // WheelBalancer::Parameters (upkie/cpp/controllers/WheelBalancer.h:48-91)
  T contact_radius, dt, fall_pitch, max_ground_velocity, pitch_damping, pitch_stiffness, position_damping,
      position_stiffness, stiff_yaw_velocity, wheel_radius;
};

template <typename T>
struct WheelBalancerState {
  T ground_velocity, integral_velocity, target_ground_position, target_yaw_velocity;
};

// constants of WheelBalancer.cpp:11-15
#define UPKIE_WB_AIR_RETURN_PERIOD 1.0
#define UPKIE_WB_MAX_INTEGRAL_VELOCITY 10.0
#define UPKIE_WB_MAX_TARGET_DISTANCE 1.0
#define UPKIE_WB_GAIN_SCALE 2.0
#define UPKIE_WB_TURNING_GAIN_SCALE 2.0

template <typename T>
UPKIE_HD T wb_clamp(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }  // std::clamp

// WheelBalancer::read (WheelBalancer.cpp:35-88)
template <typename T>
UPKIE_HD void wheel_balancer_read(const WheelBalancerParams<T>& P, WheelBalancerState<T>& s, T pitch, T ground_position,
                                  bool floor_contact, bool has_target, T target_ground_velocity, T target_yaw_velocity) {
  // without a "bullet" action key the ground-velocity target is 0 for this cycle and the yaw-velocity target keeps
  // its last value (WheelBalancer.cpp:37-42: only the local is re-initialised)
  if (has_target) s.target_yaw_velocity = target_yaw_velocity;
  else target_ground_velocity = T(0);
  const T dt = P.dt;
  if ((pitch < 0 ? -pitch : pitch) > P.fall_pitch) {
    s.ground_velocity = T(0);
    return;
  }
  const T e0 = s.target_ground_position - ground_position;
  const T e1 = T(0) - pitch;  // target pitch 0
  if (floor_contact) {
    s.integral_velocity += (P.position_stiffness * e0 + P.pitch_stiffness * e1) * dt;
    s.integral_velocity = wb_clamp(s.integral_velocity, -T(UPKIE_WB_MAX_INTEGRAL_VELOCITY), T(UPKIE_WB_MAX_INTEGRAL_VELOCITY));
    s.target_ground_position += target_ground_velocity * dt;
    s.target_ground_position = wb_clamp(s.target_ground_position, ground_position - T(UPKIE_WB_MAX_TARGET_DISTANCE),
                                        ground_position + T(UPKIE_WB_MAX_TARGET_DISTANCE));
  } else {
    // low_pass_filter(prev, cutoff, input, dt) = prev + dt / cutoff * (input - prev) (upkie/cpp/utils/low_pass_filter.h:36-37)
    const T alpha = dt / T(UPKIE_WB_AIR_RETURN_PERIOD);
    s.integral_velocity = s.integral_velocity + alpha * (T(0) - s.integral_velocity);
    s.target_ground_position = s.target_ground_position + alpha * (ground_position - s.target_ground_position);
  }
  // the error is the one computed BEFORE the target update, as in the reference
  const T trick_velocity = -target_ground_velocity;  // non-minimum phase trick
  s.ground_velocity = trick_velocity - (P.position_damping * e0 + P.pitch_damping * e1) - s.integral_velocity;
  s.ground_velocity = wb_clamp(s.ground_velocity, -P.max_ground_velocity, P.max_ground_velocity);
}

// WheelStopper::write (WheelStopper.cpp:15-22) then WheelBalancer::write (WheelBalancer.cpp:90-110) on a
// servo action a[6 joints][6 keys] (ACTION_KEYS order)
template <typename T>
UPKIE_HD void wheel_balancer_write(const WheelBalancerParams<T>& P, const WheelBalancerState<T>& s, T* a, T nan_value) {
  for (int j = 2; j < 6; j += 3) {
    a[j * 6 + UPKIE_ACT_FEEDFORWARD_TORQUE] = T(0);
    a[j * 6 + UPKIE_ACT_POSITION] = nan_value;
    a[j * 6 + UPKIE_ACT_VELOCITY] = T(0);
  }
  const T ref = s.ground_velocity / P.wheel_radius;
  a[2 * 6 + UPKIE_ACT_VELOCITY] += ref;
  a[5 * 6 + UPKIE_ACT_VELOCITY] -= ref;
  const T yaw_to_wheel = P.contact_radius / P.wheel_radius;
  a[2 * 6 + UPKIE_ACT_VELOCITY] += yaw_to_wheel * s.target_yaw_velocity;
  a[5 * 6 + UPKIE_ACT_VELOCITY] += yaw_to_wheel * s.target_yaw_velocity;
  const T ty = s.target_yaw_velocity < 0 ? -s.target_yaw_velocity : s.target_yaw_velocity;
  const T turning = ty > P.stiff_yaw_velocity ? T(1) : T(0);
  const T scale = T(UPKIE_WB_GAIN_SCALE) + T(UPKIE_WB_TURNING_GAIN_SCALE) * turning;
  const int legs[4] = {0, 1, 3, 4};
  for (int k = 0; k < 4; ++k) {
    a[legs[k] * 6 + UPKIE_ACT_KP_SCALE] = scale;
    a[legs[k] * 6 + UPKIE_ACT_KD_SCALE] = scale;
  }
}

}  // namespace upkie_b200

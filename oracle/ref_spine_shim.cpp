// SPDX-License-Identifier: Apache-2.0
//
// ref_spine_shim.cpp -- TEST INFRASTRUCTURE. C entry points around the REFERENCE'S OWN spine observers and
// controllers, compiled unmodified and in place from /root/reference (oracle/Makefile, target _ref) against the
// stand-in headers of oracle/standin/ (Eigen, palimpsest and spdlog are not in this image):
//   spines/common/observers.h:23-44      make_observers(): BaseOrientation -> FloorContact -> WheelOdometry
//   spines/common/controllers.h:24-44    make_controllers("wheel_balancer"): WheelStopper -> WheelBalancer
//   upkie/cpp/observers/*.cpp, upkie/cpp/controllers/*.cpp
// Only this file (the flat-array <-> Dictionary glue) is ours. The resulting oracle/_ref/libupkie_ref_spine.so pins
// oracle/upkie_oracle.cpp's ObserverPipelineOracle and WheelBalancerOracle on the reference's code
// (tests/test_ref_spine.py, tests/golden/make_ref_spine_golden.py). Never loaded by the product.
#include <memory>
#include <string>

#include "../include/upkie_b200.h"
#include "spines/common/controllers.h"
#include "spines/common/observers.h"

using palimpsest::Dictionary;

namespace {

const char* kJoints[6] = {"left_hip", "left_knee", "left_wheel", "right_hip", "right_knee", "right_wheel"};
const char* kActKeys[6] = {"position", "velocity", "feedforward_torque", "kp_scale", "kd_scale", "maximum_torque"};

struct RefSpine {
  upkie::cpp::observers::ObserverPipeline observers;
  upkie::cpp::controllers::ControllerPipeline controllers;
  Dictionary config;
  RefSpine(unsigned f) : observers(spines::common::make_observers(f)), controllers(spines::common::make_controllers("wheel_balancer", f)) {}
};

}  // namespace

extern "C" {

// spine_frequency: observers and controllers run one cycle per call with dt = 1 / spine_frequency
void* ref_spine_create(const UpkieObserverConfig* oc, const UpkieWheelBalancerConfig* wc, unsigned spine_frequency) {
  RefSpine* s = new RefSpine(spine_frequency);
  Dictionary& c = s->config;
  // the spine configuration dictionary (upkie/envs/backends/spine_backend.py:77-105,140-165)
  Eigen::Matrix3d Rbi;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Rbi(i, j) = oc->rotation_base_to_imu[3 * i + j];
  c("base_orientation")("rotation_base_to_imu") = Rbi;
  c("floor_contact")("upper_leg_torque_threshold") = oc->upper_leg_torque_threshold;
  c("wheel_contact")("cutoff_period") = oc->cutoff_period;
  c("wheel_contact")("liftoff_inertia") = oc->liftoff_inertia;
  c("wheel_contact")("min_touchdown_acceleration") = oc->min_touchdown_acceleration;
  c("wheel_contact")("min_touchdown_torque") = oc->min_touchdown_torque;
  c("wheel_contact")("touchdown_inertia") = oc->touchdown_inertia;
  c("wheel_odometry")("signed_radius")("left_wheel") = oc->signed_radius[0];
  c("wheel_odometry")("signed_radius")("right_wheel") = oc->signed_radius[1];
  if (wc) {
    Dictionary& w = c("wheel_balancer");
    w("contact_radius") = wc->contact_radius;
    w("fall_pitch") = wc->fall_pitch;
    w("max_ground_velocity") = wc->max_ground_velocity;
    w("pitch_damping") = wc->pitch_damping;
    w("pitch_stiffness") = wc->pitch_stiffness;
    w("position_damping") = wc->position_damping;
    w("position_stiffness") = wc->position_stiffness;
    w("stiff_yaw_velocity") = wc->stiff_yaw_velocity;
    w("wheel_radius") = wc->wheel_radius;
  }
  s->observers.reset(c);
  s->controllers.reset(c);
  return s;
}

void ref_spine_destroy(void* h) { delete static_cast<RefSpine*>(h); }

void ref_spine_reset(void* h) {
  RefSpine* s = static_cast<RefSpine*>(h);
  s->observers.reset(s->config);
  s->controllers.reset(s->config);
}

// one observer cycle: spine[UPKIE_SPINE_DIM] supplies imu.orientation, imu.angular_velocity and the servo block
void ref_spine_observers_step(void* h, const double* spine, double* out) {
  RefSpine* s = static_cast<RefSpine*>(h);
  Dictionary obs;
  const double* q = spine + UPKIE_SP_IMU_QUAT;
  obs("imu")("orientation") = Eigen::Quaterniond(q[0], q[1], q[2], q[3]);
  const double* w = spine + UPKIE_SP_IMU_ANGVEL;
  obs("imu")("angular_velocity") = Eigen::Vector3d(w[0], w[1], w[2]);
  for (int j = 0; j < 6; ++j) {
    const double* so = spine + UPKIE_SP_SERVO + 5 * j;
    obs("servo")(kJoints[j])("position") = so[UPKIE_OBS_POSITION];
    obs("servo")(kJoints[j])("velocity") = so[UPKIE_OBS_VELOCITY];
    obs("servo")(kJoints[j])("torque") = so[UPKIE_OBS_TORQUE];
  }
  s->observers.run(obs);
  const Dictionary& co = obs;
  out[UPKIE_OBSV_PITCH] = co("base_orientation")("pitch").as<double>();
  const Eigen::Vector3d& av = co("base_orientation")("angular_velocity").as<Eigen::Vector3d>();
  const Eigen::Matrix3d& R = co("base_orientation")("rotation_base_to_world").as<Eigen::Matrix3d>();
  for (int i = 0; i < 3; ++i) out[UPKIE_OBSV_ANGVEL + i] = av[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[UPKIE_OBSV_ROT + 3 * i + j] = R(i, j);
  out[UPKIE_OBSV_CONTACT] = co("floor_contact")("contact").as<bool>() ? 1.0 : 0.0;
  out[UPKIE_OBSV_WHEEL_CONTACT] = co("floor_contact")("left_wheel")("contact").as<bool>() ? 1.0 : 0.0;
  out[UPKIE_OBSV_WHEEL_CONTACT + 1] = co("floor_contact")("right_wheel")("contact").as<bool>() ? 1.0 : 0.0;
  out[UPKIE_OBSV_LEG_TORQUE] = co("floor_contact")("upper_leg_torque").as<double>();
  out[UPKIE_OBSV_WHEEL_INERTIA] = co("floor_contact")("left_wheel")("inertia").as<double>();
  out[UPKIE_OBSV_WHEEL_INERTIA + 1] = co("floor_contact")("right_wheel")("inertia").as<double>();
  out[UPKIE_OBSV_ODOM_POS] = co("wheel_odometry")("position").as<double>();
  out[UPKIE_OBSV_ODOM_VEL] = co("wheel_odometry")("velocity").as<double>();
}

// one controller cycle: obs3 = pitch, floor contact (0 / 1), wheel-odometry position; target2 = the "bullet" action
// key (target ground / yaw velocity) or NULL; action[6][6] in ACTION_KEYS order, updated in place
void ref_spine_controllers_step(void* h, const double* obs3, const double* target2, double* action) {
  RefSpine* s = static_cast<RefSpine*>(h);
  Dictionary obs, act;
  obs("base_orientation")("pitch") = obs3[0];
  obs("floor_contact")("contact") = (obs3[1] != 0.0);
  obs("wheel_odometry")("position") = obs3[2];
  if (target2) {
    act("bullet")("target_ground_velocity") = target2[0];
    act("bullet")("target_yaw_velocity") = target2[1];
  }
  for (int j = 0; j < 6; ++j)
    for (int k = 0; k < 6; ++k) act("servo")(kJoints[j])(kActKeys[k]) = action[6 * j + k];
  s->controllers.run(obs, act);
  const Dictionary& ca = act;
  for (int j = 0; j < 6; ++j)
    for (int k = 0; k < 6; ++k) action[6 * j + k] = ca("servo")(kJoints[j])(kActKeys[k]).as<double>();
}

}  // extern "C"

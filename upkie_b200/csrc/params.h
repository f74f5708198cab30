// SPDX-License-Identifier: Apache-2.0
//
// params.h -- host-side conversion of the public UpkieModel / UpkieSimConfig
// (double precision, include/upkie_b200.h) into the fp32 kernel parameter block.
#pragma once

#include <cmath>
#include <string>

#include "sim_core.cuh"

namespace upkie_b200 {

inline void default_sim_config(UpkieSimConfig* c) {
  // PyBulletBackend.__init__ defaults (upkie/envs/backends/pybullet_backend.py:55-112)
  c->dt = 1.0 / 200.0;          // UpkieServos frequency=200 (upkie/envs/upkie_servos.py:118)
  c->nb_substeps = 5;           // int(1000 * dt), pybullet_backend.py:85-87
  c->pgs_iterations = 50;       // Bullet numSolverIterations default
  c->gravity = 9.81;            // pybullet_backend.py:110
  c->torque_control_kp = 20.0;  // pybullet_backend.py:65
  c->torque_control_kd = 1.0;   // pybullet_backend.py:64
  for (int j = 0; j < UPKIE_NJ; ++j) {
    c->joint_friction[j] = 0.0;
    c->torque_control_noise[j] = 0.0;
    c->torque_measurement_noise[j] = 0.0;
  }
  for (int k = 0; k < 3; ++k) c->imu_accelerometer_bias[k] = c->imu_gyroscope_bias[k] = 0.0;
  c->imu_accelerometer_noise = c->imu_gyroscope_noise = 0.0;
  c->noise_seed = 0;
  c->linear_damping = 0.04;     // Bullet btMultiBody default
  c->angular_damping = 0.04;
  c->max_coordinate_velocity = 100.0;
  c->contact_stiffness = 30000.0;
  c->contact_damping = 1000.0;
  c->contact_breaking_threshold = 0.02;
  c->friction = 1.0;            // plane.urdf lateral_friction=1 (upkie/cpp/interfaces/bullet/plane/plane.urdf:5) x tire
  c->max_gain_scale = 5.0;      // upkie_servos.py:120
  c->fall_pitch = 1.0;          // upkie_gyropod.py:108
  c->leg_gain_scale = 1.0;
  c->max_ground_velocity = 3.0;
  c->max_yaw_velocity = 1.0;
  c->servos_fall_termination = 0;
  c->skip_action_clamps = 0;
  c->min_base_height = 0.0;
  c->pgs_tolerance = 0.0;  // deprecated, ignored (see solver_residual_threshold)
  c->warmstarting_factor = 0.0;  // off by default (Bullet's m_warmstartingFactor is 0.85)
  c->joint_limits = 3;  // Bullet's hip / knee limit constraints ON (pybullet_backend.py:121 loadURDF): ten-row solver for warps with a robot on a bound
  c->reserved_joint_limits = 0;
  c->joint_limit_erp = 0.2;
  c->joint_limit_max_impulse = 100.0;
  c->init_position[0] = 0.0; c->init_position[1] = 0.0; c->init_position[2] = 0.6;  // upkie_env.py:87-90
  c->init_quat[0] = 1.0; c->init_quat[1] = 0.0; c->init_quat[2] = 0.0; c->init_quat[3] = 0.0;
  c->rand_roll = c->rand_pitch = c->rand_x = c->rand_z = 0.0;
  c->rand_omega_x = c->rand_omega_y = 0.0;
  for (int k = 0; k < 3; ++k) c->rand_linear_velocity[k] = 0.0;
  for (int j = 0; j < UPKIE_NJ; ++j) c->init_joint_configuration[j] = 0.0;
  for (int k = 0; k < 3; ++k) c->init_angular_velocity[k] = c->init_linear_velocity[k] = 0.0;
  c->spine_mode = 0;
  c->reserved_spine_mode = 0;
  c->body_contacts = 0;  // opt-in for batched handles (B200Backend, the single-env drop-in, turns it on): see include/upkie_b200.h
  c->reserved_body_contacts = 0;
  c->body_contact_erp = 0.2;  // btContactSolverInfo::m_erp2
  c->body_friction = 0.5;     // URDF importer default lateral friction of a link without <contact>
  c->solver_residual_threshold = 1e-7;  // PyBullet: getSolverInfo().m_leastSquaresResidualThreshold = 1e-7
}

inline void default_mpc_config(UpkieMpcConfig* c) {
  // MPCBalancer.__init__ defaults (upkie/controllers/mpc_balancer.py:168-181)
  c->fall_pitch = 1.0;
  c->leg_length = 0.58;
  c->max_ground_accel = 10.0;
  c->max_ground_velocity = 3.0;
  c->nb_timesteps = 50;
  c->max_iterations = 30;
  c->sampling_period = 0.02;
  c->stage_input_cost_weight = 1e-3;
  c->stage_state_cost_weight = 1e-3;
  c->terminal_cost_weight = 1.0;
  c->gravity = 9.81;
}

// Returns 0 on success; fills `err` otherwise.
inline int make_sim_params(const UpkieModel& m, const UpkieSimConfig& c, SimParams& P, std::string& err) {
  const int expect_parent[UPKIE_NB] = {-1, 0, 1, 2, 0, 4, 5};
  for (int i = 0; i < UPKIE_NB; ++i)
    if (m.parent[i] != expect_parent[i]) {
      err = "model: kinematic tree must be base -> (hip, knee, wheel) x 2 in URDF joint order";
      return UPKIE_B200_EMODEL;
    }
  P.any_ctrl_noise = 0;
  P.any_meas_noise = 0;
  for (int j = 0; j < UPKIE_NJ; ++j) {
    const double ax = m.joint_axis[j][0], ay = m.joint_axis[j][1], az = m.joint_axis[j][2];
    if (std::fabs(ax) > 1e-9 || std::fabs(az) > 1e-9 || std::fabs(std::fabs(ay) - 1.0) > 1e-9) {
      err = "model: the sm_100a kernels specialise on joint axes along +-y of the base frame (Upkie, Cookie)";
      return UPKIE_B200_EMODEL;
    }
    P.sgn[j] = ay > 0 ? 1.f : -1.f;
    for (int k = 0; k < 3; ++k) P.jo[j][k] = float(m.joint_origin[j][k]);
    P.q_lower[j] = float(m.q_lower[j]);
    P.q_upper[j] = float(m.q_upper[j]);
    P.qd_max[j] = float(m.qd_max[j]);
    P.tau_max[j] = float(m.tau_max[j]);
    P.joint_friction[j] = float(c.joint_friction[j]);
    P.ctrl_noise[j] = float(c.torque_control_noise[j]);
    P.meas_noise[j] = float(c.torque_measurement_noise[j]);
    // thresholds of the reference: noise is drawn only when sigma > 1e-10 (pybullet_backend.py:464,548)
    if (c.torque_control_noise[j] > 1e-10) P.any_ctrl_noise = 1;
    if (c.torque_measurement_noise[j] > 1e-10) P.any_meas_noise = 1;
  }
  for (int i = 0; i < UPKIE_NB; ++i) {
    if (!(m.mass[i] > 0.0)) { err = "model: body masses must be positive"; return UPKIE_B200_EMODEL; }
    P.mass[i] = float(m.mass[i]);
    for (int k = 0; k < 3; ++k) P.com[i][k] = float(m.com[i][k]);
    for (int k = 0; k < 6; ++k) P.inertia[i][k] = float(m.inertia[i][k]);
  }
  auto wheel_sym = [&](int b) {
    const double* I = m.inertia[b];
    const double* cm = m.com[b];
    return std::fabs(cm[0]) < 1e-12 && std::fabs(cm[2]) < 1e-12 && std::fabs(I[0] - I[2]) < 1e-12 &&
           std::fabs(I[3]) < 1e-12 && std::fabs(I[4]) < 1e-12 && std::fabs(I[5]) < 1e-12;
  };
  P.wheel_symmetric = (wheel_sym(3) && wheel_sym(6)) ? 1 : 0;
  P.noise_seed = c.noise_seed;
  P.any_imu_uncertainty = 0;
  for (int k = 0; k < 3; ++k) {
    P.imu_acc_bias[k] = float(c.imu_accelerometer_bias[k]);
    P.imu_gyro_bias[k] = float(c.imu_gyroscope_bias[k]);
    if (c.imu_accelerometer_bias[k] != 0.0 || c.imu_gyroscope_bias[k] != 0.0) P.any_imu_uncertainty = 1;
  }
  P.imu_acc_noise = float(c.imu_accelerometer_noise);
  P.imu_gyro_noise = float(c.imu_gyroscope_noise);
  if (c.imu_accelerometer_noise > 0.0 || c.imu_gyroscope_noise > 0.0) P.any_imu_uncertainty = 1;
  for (int k = 0; k < 3; ++k) {
    P.sgn2[k].x = P.sgn[k]; P.sgn2[k].y = P.sgn[k + 3];
    P.mass2[k].x = P.mass[k + 1]; P.mass2[k].y = P.mass[k + 4];
    float oyl = 0.f, oyr = 0.f;
    for (int kk = 0; kk <= k; ++kk) { oyl += P.jo[kk][1]; oyr += P.jo[kk + 3][1]; }
    P.oy2[k].x = oyl; P.oy2[k].y = oyr;
    for (int i = 0; i < 3; ++i) {
      P.jo2[k][i].x = P.jo[k][i]; P.jo2[k][i].y = P.jo[k + 3][i];
      P.com2[k][i].x = P.com[k + 1][i]; P.com2[k][i].y = P.com[k + 4][i];
    }
    for (int i = 0; i < 6; ++i) { P.inertia2[k][i].x = P.inertia[k + 1][i]; P.inertia2[k][i].y = P.inertia[k + 4][i]; }
  }
  if (!(m.wheel_radius > 0.0)) { err = "model: wheel_radius must be positive"; return UPKIE_B200_EMODEL; }
  P.wheel_radius = float(m.wheel_radius);
  P.half_wheel_base = float(0.5 * m.wheel_base);
  P.left_sign = m.left_wheeled ? 1.f : -1.f;
  for (int k = 0; k < 3; ++k) P.imu_pos[k] = float(m.imu_position[k]);
  for (int k = 0; k < 9; ++k) P.Rbi[k] = float(m.rotation_base_to_imu[k]);

  if (!(c.dt > 0.0) || c.nb_substeps < 1 || c.pgs_iterations < 0) {
    err = "config: dt > 0, nb_substeps >= 1, pgs_iterations >= 0 required";
    return UPKIE_B200_EINVAL;
  }
  const double h = c.dt / c.nb_substeps;
  P.dt = float(c.dt);
  P.inv_dt = float(1.0 / c.dt);
  P.h = float(h);
  P.inv_h = float(1.0 / h);
  P.nb_substeps = c.nb_substeps;
  P.pgs_iterations = c.pgs_iterations;
  P.res_thr = float(c.solver_residual_threshold);
  P.warm = float(c.warmstarting_factor);
  P.joint_limits = c.joint_limits < 0 ? 0 : (c.joint_limits > 3 ? 3 : c.joint_limits);
  P.limit_erp = float(c.joint_limit_erp);
  P.limit_max_impulse = float(c.joint_limit_max_impulse);
  P.skip_action_clamps = c.skip_action_clamps;
  P.gravity = float(c.gravity);
  P.kp = float(c.torque_control_kp);
  P.kd = float(c.torque_control_kd);
  P.lin_damp = float(c.linear_damping);
  P.ang_damp = float(c.angular_damping);
  P.vmax = float(c.max_coordinate_velocity);
  // Bullet btMultiBodyConstraintSolver: cfm = 1/(h k + d), erp = h k/(h k + d), cfm *= 1/h
  double denom = h * c.contact_stiffness + c.contact_damping;
  if (denom < 1.1920929e-7) denom = 1.1920929e-7;
  P.cfm = float((1.0 / denom) / h);
  P.erp = float(h * c.contact_stiffness / denom);
  P.breaking_threshold = float(c.contact_breaking_threshold);
  P.friction = float(c.friction);
  P.max_gain_scale = float(c.max_gain_scale);
  P.fall_pitch = float(c.fall_pitch);
  P.leg_gain_scale = float(c.leg_gain_scale);
  P.max_ground_velocity = float(c.max_ground_velocity);
  P.max_yaw_velocity = float(c.max_yaw_velocity);
  P.servos_fall_termination = c.servos_fall_termination;
  P.min_base_height = float(c.min_base_height);
  for (int k = 0; k < 3; ++k) P.init_pos[k] = float(c.init_position[k]);
  for (int k = 0; k < 4; ++k) P.init_quat[k] = float(c.init_quat[k]);
  P.rand_roll = float(c.rand_roll);
  P.rand_pitch = float(c.rand_pitch);
  P.rand_x = float(c.rand_x);
  P.rand_z = float(c.rand_z);
  P.rand_omega_x = float(c.rand_omega_x);
  P.rand_omega_y = float(c.rand_omega_y);
  for (int k = 0; k < 3; ++k) P.rand_linvel[k] = float(c.rand_linear_velocity[k]);
  P.spine_mode = c.spine_mode ? 1 : 0;
  if (P.spine_mode && P.joint_limits == 0) {
    err = "config: spine_mode needs joint_limits != 0 (it lives in the extras + limits kernels)";
    return UPKIE_B200_EINVAL;
  }
  // body-ground contacts: the model's collision points and the gate constants of the packed solvers
  if (m.n_collision_points < 0 || m.n_collision_points > UPKIE_MAX_COLLISION_POINTS) {
    err = "model: n_collision_points must be in [0, UPKIE_MAX_COLLISION_POINTS]";
    return UPKIE_B200_EMODEL;
  }
  P.n_bp = m.n_collision_points;
  P.n_gate_base = 0;
  for (int k = 0; k < 3; ++k) { P.gate_leg_bound[k].x = -1e30f; P.gate_leg_bound[k].y = -1e30f; }
  for (int p = 0; p < UPKIE_MAX_COLLISION_POINTS; ++p) {
    P.bp_body[p] = 0; P.bp_radius[p] = 0.f;
    for (int k = 0; k < 3; ++k) P.bp_pos[p][k] = 0.f;
    for (int k = 0; k < 4; ++k) P.gate_base[p][k] = 0.f;
  }
  for (int p = 0; p < P.n_bp; ++p) {
    const int b = m.collision_body[p];
    if (b < 0 || b >= UPKIE_NB || !(m.collision_radius[p] >= 0.0)) {
      err = "model: collision point on an unknown body or with a negative radius";
      return UPKIE_B200_EMODEL;
    }
    P.bp_body[p] = b;
    for (int k = 0; k < 3; ++k) P.bp_pos[p][k] = float(m.collision_point[p][k]);
    P.bp_radius[p] = float(m.collision_radius[p]);
    const double* cp = m.collision_point[p];
    if (b == 0) {
      float* g = P.gate_base[P.n_gate_base++];
      g[0] = P.bp_pos[p][0]; g[1] = P.bp_pos[p][1]; g[2] = P.bp_pos[p][2]; g[3] = P.bp_radius[p];
    } else {
      // a point on a leg body can be at most |point| + radius below that body's origin (1 mm of slack for fp32)
      const float bound = float(std::sqrt(cp[0] * cp[0] + cp[1] * cp[1] + cp[2] * cp[2]) + m.collision_radius[p] + 1e-3);
      const int leg = (b - 1) / 3, k = (b - 1) % 3;
      float& slot = leg == 0 ? P.gate_leg_bound[k].x : P.gate_leg_bound[k].y;
      if (bound > slot) slot = bound;
    }
  }
  P.body_contacts = (c.body_contacts != 0 && P.n_bp > 0 && P.joint_limits != 0) ? 1 : 0;
  P.body_erp = float(c.body_contact_erp);
  P.body_mu_scale = float(c.body_friction);
  for (int j = 0; j < 6; ++j) P.init_q[j] = float(c.init_joint_configuration[j]);
  for (int k = 0; k < 3; ++k) {
    P.init_angvel[k] = float(c.init_angular_velocity[k]);
    P.init_linvel[k] = float(c.init_linear_velocity[k]);
  }
  return 0;
}

// AoS state row <-> registers
UPKIE_HD void state_from_row(const float* r, RobotState& S) {
  for (int i = 0; i < 3; ++i) {
    S.pos[i] = r[UPKIE_ST_POS + i];
    S.linvel[i] = r[UPKIE_ST_LINVEL + i];
    S.angvel[i] = r[UPKIE_ST_ANGVEL + i];
    S.prev_imu_vel[i] = r[UPKIE_ST_PREV_IMU_VEL + i];
    S.imu_acc[i] = r[UPKIE_ST_IMU_ACC + i];
  }
  for (int i = 0; i < 4; ++i) {
    S.quat[i] = r[UPKIE_ST_QUAT + i];
    S.leg_target[i] = r[UPKIE_ST_LEG_TARGET + i];
  }
  for (int j = 0; j < 6; ++j) {
    S.q[j] = r[UPKIE_ST_Q + j];
    S.qd[j] = r[UPKIE_ST_QD + j];
    S.torque[j] = r[UPKIE_ST_TORQUE + j];
  }
  S.yaw = r[UPKIE_ST_YAW];
  S.yaw_vel = r[UPKIE_ST_YAW_VEL];
  S.contact = r[UPKIE_ST_CONTACT];
  S.lam_n[0] = r[UPKIE_ST_CONTACT_IMPULSE];
  S.lam_n[1] = r[UPKIE_ST_CONTACT_IMPULSE + 1];
  for (int k = 0; k < 4; ++k) S.lam_t[k] = r[UPKIE_ST_FRICTION_IMPULSE + k];
}

UPKIE_HD void state_to_row(const RobotState& S, float* r) {
  for (int i = 0; i < 3; ++i) {
    r[UPKIE_ST_POS + i] = S.pos[i];
    r[UPKIE_ST_LINVEL + i] = S.linvel[i];
    r[UPKIE_ST_ANGVEL + i] = S.angvel[i];
    r[UPKIE_ST_PREV_IMU_VEL + i] = S.prev_imu_vel[i];
    r[UPKIE_ST_IMU_ACC + i] = S.imu_acc[i];
  }
  for (int i = 0; i < 4; ++i) {
    r[UPKIE_ST_QUAT + i] = S.quat[i];
    r[UPKIE_ST_LEG_TARGET + i] = S.leg_target[i];
  }
  for (int j = 0; j < 6; ++j) {
    r[UPKIE_ST_Q + j] = S.q[j];
    r[UPKIE_ST_QD + j] = S.qd[j];
    r[UPKIE_ST_TORQUE + j] = S.torque[j];
  }
  r[UPKIE_ST_YAW] = S.yaw;
  r[UPKIE_ST_YAW_VEL] = S.yaw_vel;
  r[UPKIE_ST_CONTACT] = S.contact;
  r[UPKIE_ST_CONTACT_IMPULSE] = S.lam_n[0];
  r[UPKIE_ST_CONTACT_IMPULSE + 1] = S.lam_n[1];
  for (int k = 0; k < 4; ++k) r[UPKIE_ST_FRICTION_IMPULSE + k] = S.lam_t[k];
}

}  // namespace upkie_b200

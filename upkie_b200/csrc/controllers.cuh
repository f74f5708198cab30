// SPDX-License-Identifier: Apache-2.0
//
// controllers.cuh -- batched "wheel_balancer" controller pipeline: kernel + handle management.
// One thread = one robot; controller state is struct-of-arrays [4][n].
#pragma once

#include <cuda_runtime.h>

#include <new>
#include <string>

#include "controllers_core.cuh"

namespace upkie_b200 {

struct WheelBalancerHandle {
  uint32_t magic;
  int n, device;
  WheelBalancerParams<float> P;
  float* state = nullptr;  // [4][n]
};
constexpr uint32_t kWheelBalancerMagic = 0x55505742u;

inline void default_wheel_balancer_config(UpkieWheelBalancerConfig* c) {
  // WheelBalancer::Parameters defaults (WheelBalancer.h:60-90) with the overrides of the spine
  // (spines/common/controllers.h:33-36: dt = 1 / spine_frequency, wheel_radius = 0.06)
  c->contact_radius = 0.1524;
  c->dt = 1.0 / 1000.0;
  c->fall_pitch = 1.0;
  c->max_ground_velocity = 2.0;
  c->pitch_damping = 1.8;
  c->pitch_stiffness = 20.0;
  c->position_damping = 0.7;
  c->position_stiffness = 1.6;
  c->stiff_yaw_velocity = 0.1;
  c->wheel_radius = 0.06;
}

__global__ void k_wheel_balancer_step(const __grid_constant__ WheelBalancerParams<float> P, int n,
                                      float* __restrict__ state, const float* __restrict__ obs, int obs_stride,
                                      int pitch_off, int contact_off, int odom_off, const float* __restrict__ target,
                                      float* __restrict__ action) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  WheelBalancerState<float> s{state[i], state[size_t(n) + i], state[size_t(2) * n + i], state[size_t(3) * n + i]};
  const float* o = obs + size_t(i) * obs_stride;
  const float tgv = target ? target[2 * i] : 0.f, tyv = target ? target[2 * i + 1] : 0.f;
  wheel_balancer_read(P, s, o[pitch_off], o[odom_off], o[contact_off] != 0.f, target != nullptr, tgv, tyv);
  float a[UPKIE_ACT_DIM];
  for (int k = 0; k < UPKIE_ACT_DIM; ++k) a[k] = action[size_t(i) * UPKIE_ACT_DIM + k];
  wheel_balancer_write(P, s, a, __int_as_float(0x7fc00000));
  for (int k = 0; k < UPKIE_ACT_DIM; ++k) action[size_t(i) * UPKIE_ACT_DIM + k] = a[k];
  state[i] = s.ground_velocity;
  state[size_t(n) + i] = s.integral_velocity;
  state[size_t(2) * n + i] = s.target_ground_position;
  state[size_t(3) * n + i] = s.target_yaw_velocity;
}

__global__ void k_wheel_balancer_reset(int n, const uint8_t* __restrict__ mask, float* __restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && !mask[i]) return;
  for (int k = 0; k < 4; ++k) state[size_t(k) * n + i] = 0.f;
}

inline WheelBalancerHandle* as_wheel_balancer(void* p) {
  WheelBalancerHandle* h = static_cast<WheelBalancerHandle*>(p);
  return (h && h->magic == kWheelBalancerMagic) ? h : nullptr;
}

inline void wheel_balancer_destroy_impl(void* p) {
  WheelBalancerHandle* h = as_wheel_balancer(p);
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->state);
  h->magic = 0;
  delete h;
}

inline int wheel_balancer_create_impl(const UpkieWheelBalancerConfig& c, int n, int device, void** out, std::string& err) {
  if (n < 1) { err = "wheel_balancer_create: n_robots must be >= 1"; return UPKIE_B200_EINVAL; }
  // low_pass_filter throws FilterError when the cutoff period (1 s air return) <= 2 dt (low_pass_filter.h:26-33)
  if (!(c.dt > 0.0) || UPKIE_WB_AIR_RETURN_PERIOD <= 2.0 * c.dt || !(c.wheel_radius > 0.0)) {
    err = "wheel_balancer_create: need 0 < dt < 0.5 s and wheel_radius > 0";
    return UPKIE_B200_EINVAL;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    err = "wheel_balancer_create: no CUDA device available (this library has no CPU path)";
    return UPKIE_B200_ECUDA;
  }
  if (device < 0 || device >= count) { err = "wheel_balancer_create: invalid device index"; return UPKIE_B200_EINVAL; }
  WheelBalancerHandle* h = new (std::nothrow) WheelBalancerHandle();
  if (!h) { err = "wheel_balancer_create: out of host memory"; return UPKIE_B200_ENOMEM; }
  h->P = WheelBalancerParams<float>{float(c.contact_radius), float(c.dt), float(c.fall_pitch), float(c.max_ground_velocity),
                                    float(c.pitch_damping), float(c.pitch_stiffness), float(c.position_damping),
                                    float(c.position_stiffness), float(c.stiff_yaw_velocity), float(c.wheel_radius)};
  h->magic = kWheelBalancerMagic;
  h->n = n;
  h->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&h->state, size_t(4) * n * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(h->state, 0, size_t(4) * n * sizeof(float));
  if (e != cudaSuccess) {
    err = std::string("wheel_balancer_create: ") + cudaGetErrorString(e);
    wheel_balancer_destroy_impl(h);
    return UPKIE_B200_ECUDA;
  }
  *out = h;
  return UPKIE_B200_OK;
}

}  // namespace upkie_b200

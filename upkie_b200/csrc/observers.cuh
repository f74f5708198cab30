// SPDX-License-Identifier: Apache-2.0
//
// observers.cuh -- batched spine observer pipeline: kernel + handle management.
// One thread = one robot; observer state is struct-of-arrays [13][n] (coalesced).
#pragma once

#include <cuda_runtime.h>

#include <new>
#include <string>

#include "observers_core.cuh"

namespace upkie_b200 {

constexpr int kObserverStateDim = 13;

struct ObserversHandle {
  uint32_t magic;
  int n, device;
  ObserverParams<float> P;
  float* state = nullptr;  // [13][n]
};
constexpr uint32_t kObserversMagic = 0x55504F42u;

inline int make_observer_params(const UpkieObserverConfig& c, ObserverParams<float>& P, std::string& err) {
  // low_pass_filter throws FilterError when cutoff_period <= 2 dt (upkie/cpp/utils/low_pass_filter.h:27-34)
  if (!(c.dt > 0.0) || c.cutoff_period <= 2.0 * c.dt || 0.01 <= 2.0 * c.dt) {
    err = "observers: dt must be positive and both cutoff periods (wheel contact, 0.01 s leg torque) > 2 dt";
    return UPKIE_B200_EINVAL;
  }
  P.dt = float(c.dt);
  P.cutoff_period = float(c.cutoff_period);
  P.liftoff_inertia = float(c.liftoff_inertia);
  P.min_touchdown_acceleration = float(c.min_touchdown_acceleration);
  P.min_touchdown_torque = float(c.min_touchdown_torque);
  P.touchdown_inertia = float(c.touchdown_inertia);
  P.upper_leg_torque_threshold = float(c.upper_leg_torque_threshold);
  P.signed_radius[0] = float(c.signed_radius[0]);
  P.signed_radius[1] = float(c.signed_radius[1]);
  for (int k = 0; k < 9; ++k) P.Rbi[k] = float(c.rotation_base_to_imu[k]);
  return 0;
}

inline void default_observer_config(const UpkieModel& m, UpkieObserverConfig* c) {
  // _DEFAULT_SPINE_CONFIG and the model-derived entries (upkie/envs/backends/spine_backend.py:77-105,140-165)
  c->dt = 1.0 / 1000.0;  // spine frequency 1 kHz (upkie/cpp/spine/Spine.h:60)
  c->cutoff_period = 0.2;
  c->liftoff_inertia = 1e-3;
  c->min_touchdown_acceleration = 2.0;
  c->min_touchdown_torque = 0.015;
  c->touchdown_inertia = 4e-3;
  c->upper_leg_torque_threshold = 10.0;
  const double sign = m.left_wheeled ? 1.0 : -1.0;
  c->signed_radius[0] = sign * m.wheel_radius;
  c->signed_radius[1] = -sign * m.wheel_radius;
  for (int k = 0; k < 9; ++k) c->rotation_base_to_imu[k] = m.rotation_base_to_imu[k];
}

__device__ __forceinline__ void obs_load(const float* st, int n, int i, ObserverState<float>& s) {
  for (int k = 0; k < 2; ++k) {
    s.wheel[k].velocity = st[size_t(5 * k + 0) * n + i];
    s.wheel[k].abs_acceleration = st[size_t(5 * k + 1) * n + i];
    s.wheel[k].abs_torque = st[size_t(5 * k + 2) * n + i];
    s.wheel[k].inertia = st[size_t(5 * k + 3) * n + i];
    s.wheel[k].contact = st[size_t(5 * k + 4) * n + i];
  }
  s.upper_leg_torque = st[size_t(10) * n + i];
  s.odom_position = st[size_t(11) * n + i];
  s.odom_velocity = st[size_t(12) * n + i];
}

__device__ __forceinline__ void obs_store(float* st, int n, int i, const ObserverState<float>& s) {
  for (int k = 0; k < 2; ++k) {
    st[size_t(5 * k + 0) * n + i] = s.wheel[k].velocity;
    st[size_t(5 * k + 1) * n + i] = s.wheel[k].abs_acceleration;
    st[size_t(5 * k + 2) * n + i] = s.wheel[k].abs_torque;
    st[size_t(5 * k + 3) * n + i] = s.wheel[k].inertia;
    st[size_t(5 * k + 4) * n + i] = s.wheel[k].contact;
  }
  st[size_t(10) * n + i] = s.upper_leg_torque;
  st[size_t(11) * n + i] = s.odom_position;
  st[size_t(12) * n + i] = s.odom_velocity;
}

__global__ void k_observers_step(const __grid_constant__ ObserverParams<float> P, int n, float* __restrict__ state,
                                 const float* __restrict__ spine_obs, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ObserverState<float> s;
  obs_load(state, n, i, s);
  float o[UPKIE_OBSV_DIM];
  observers_step(P, s, spine_obs + size_t(i) * UPKIE_SPINE_DIM, o);
  obs_store(state, n, i, s);
  for (int k = 0; k < UPKIE_OBSV_DIM; ++k) out[size_t(i) * UPKIE_OBSV_DIM + k] = o[k];
}

__global__ void k_observers_reset(int n, const uint8_t* __restrict__ mask, float* __restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && !mask[i]) return;
  for (int k = 0; k < kObserverStateDim; ++k) state[size_t(k) * n + i] = 0.f;
}

inline ObserversHandle* as_observers(void* p) {
  ObserversHandle* h = static_cast<ObserversHandle*>(p);
  return (h && h->magic == kObserversMagic) ? h : nullptr;
}

inline void observers_destroy_impl(void* p) {
  ObserversHandle* h = as_observers(p);
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->state);
  h->magic = 0;
  delete h;
}

inline int observers_create_impl(const UpkieObserverConfig& c, int n, int device, void** out, std::string& err) {
  if (n < 1) { err = "observers_create: n_robots must be >= 1"; return UPKIE_B200_EINVAL; }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    err = "observers_create: no CUDA device available (this library has no CPU path)";
    return UPKIE_B200_ECUDA;
  }
  if (device < 0 || device >= count) { err = "observers_create: invalid device index"; return UPKIE_B200_EINVAL; }
  ObserversHandle* h = new (std::nothrow) ObserversHandle();
  if (!h) { err = "observers_create: out of host memory"; return UPKIE_B200_ENOMEM; }
  int rc = make_observer_params(c, h->P, err);
  if (rc) { delete h; return rc; }
  h->magic = kObserversMagic;
  h->n = n;
  h->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&h->state, size_t(kObserverStateDim) * n * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(h->state, 0, size_t(kObserverStateDim) * n * sizeof(float));
  if (e != cudaSuccess) {
    err = std::string("observers_create: ") + cudaGetErrorString(e);
    observers_destroy_impl(h);
    return UPKIE_B200_ECUDA;
  }
  *out = h;
  return 0;
}

inline int observers_reset_impl(void* p, const uint8_t* mask, cudaStream_t s, std::string& err) {
  ObserversHandle* h = as_observers(p);
  if (!h) { err = "invalid observers handle"; return UPKIE_B200_EINVAL; }
  cudaSetDevice(h->device);
  k_observers_reset<<<(h->n + 127) / 128, 128, 0, s>>>(h->n, mask, h->state);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { err = cudaGetErrorString(e); return UPKIE_B200_ECUDA; }
  return 0;
}

inline int observers_step_impl(void* p, const float* spine_obs, float* out, cudaStream_t s, std::string& err) {
  ObserversHandle* h = as_observers(p);
  if (!h || !spine_obs || !out) { err = "observers_step: invalid argument"; return UPKIE_B200_EINVAL; }
  cudaSetDevice(h->device);
  k_observers_step<<<(h->n + 127) / 128, 128, 0, s>>>(h->P, h->n, h->state, spine_obs, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { err = cudaGetErrorString(e); return UPKIE_B200_ECUDA; }
  return 0;
}

}  // namespace upkie_b200

// SPDX-License-Identifier: Apache-2.0
//
// upkie_b200.cu -- sm_100a kernels and the C ABI of include/upkie_b200.h.
//
// Kernels (one thread = one robot, state struct-of-arrays, model in the kernel
// parameter constant bank):
//   k_step<MODE>   one 5 ms env tick: action front-end, 5 x (moteus torque law,
//                  articulated-body dynamics, wheel-ground contact solve,
//                  semi-implicit integration), observation, termination; optional
//                  fused auto-reset.
//   k_reset        masked reset: set state, one physics substep, observe.
//   k_spine_obs / k_reset_obs / k_get_state / k_set_state   layout helpers.
// MPC kernels live in mpc.cuh.
//
// There is deliberately NO CPU path in this library: every entry point needs a
// CUDA device and fails with UPKIE_B200_ECUDA otherwise.

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "mpc.cuh"
#include "controllers.cuh"
#include "observers.cuh"
#include "kernel_common.cuh"
#include "params.h"

using namespace upkie_b200;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return fail(UPKIE_B200_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct Handle {
  uint32_t magic;
  int n, n_pad, device;
  SimParams P;
  UpkieModel model;            // kept for upkie_b200_set_config
  float* state = nullptr;      // [STATE_DIM][n_pad]
  float* eps = nullptr;        // [n][6] or null
  float* mu = nullptr;         // [n] or null
  uint32_t* err = nullptr;     // [n]
  uint8_t* done_prev = nullptr;  // [n]
  uint32_t* episode = nullptr;   // [n]
  uint32_t* tick = nullptr;      // [n] env ticks since create: counter of the noise generator
  float* ext = nullptr;          // [7 * 3][n_pad] external forces, null = none
  float* lag = nullptr;          // [UPKIE_LAG_DIM][n_pad] spine-mode lag records (config.spine_mode), else null
  float* body_rec = nullptr;     // [UPKIE_BODY_REC_DIM][n_pad] body-ground contacts of the last substep (config.body_contacts)
  uint32_t ext_local = 0;
  int autoreset = AUTORESET_DISABLED;
  uint64_t seed = 0, env_offset = 0;
  int block = UPKIE_DEFAULT_BLOCK;
  int num_sms = 148;
  int host_chunks = 2;           // chunks of the pipelined host-buffer step (measured best of 2..16, tools/e2e_parts.py)
  uint64_t step_launches = 0;    // step kernels launched (upkie_b200_launch_count)
  int host_block = 128;          // zero-copy step: persistent blocks of 4 warps, one per SM: ~3.5 tiles per block at
  int host_blocks_per_sm = 1;    // 65536 envs, so PCIe reads of tile k+1 overlap the compute of tile k
  int zero_copy = 2;             // pinned host buffers: 0 staged copies, 1 kernel reads+writes host memory, 2 hybrid
  // host-buffer staging (allocated on first use)
  float *h_act = nullptr, *h_obs = nullptr, *h_rew = nullptr;
  uint8_t *h_term = nullptr, *h_trunc = nullptr;
  float *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr;
  uint8_t *d_term = nullptr, *d_trunc = nullptr;
  cudaStream_t host_streams[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t host_events[64] = {};
  int host_kernel_streams = 1;
  double host_split[8] = {};
  int host_split_n = 0;   // hybrid host step: kernels of successive chunks alternate over this many streams
};
constexpr int kHostStreams = 3;  // H2D, kernel and D2H of different chunks overlap
constexpr uint32_t kMagic = 0x55504B42u;  // "UPKB"

Handle* as_handle(void* h) {
  Handle* p = static_cast<Handle*>(h);
  return (p && p->magic == kMagic) ? p : nullptr;
}

// ---- masked reset ------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_reset(const __grid_constant__ SimParams P, int n, int n_pad, float* __restrict__ state,
        const uint8_t* __restrict__ mask, const float* __restrict__ init_state, const float* __restrict__ eps_all,
        const float* __restrict__ mu_all, uint32_t* __restrict__ err, uint8_t* __restrict__ done_prev,
        uint32_t* __restrict__ episode, uint64_t seed, uint64_t env_offset, float* __restrict__ lag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && !mask[i]) return;
  RobotState S;
  load_state(state, n_pad, i, S);
  float epsv[6];
  const float* eps = nullptr;
  if (eps_all) {
#pragma unroll
    for (int k = 0; k < 6; ++k) epsv[k] = eps_all[size_t(i) * 6 + k];
    eps = epsv;
  }
  const float mu = mu_all ? mu_all[i] : P.friction;
  float init[UPKIE_INIT_DIM];
  if (init_state) {
#pragma unroll
    for (int k = 0; k < UPKIE_INIT_DIM; ++k) init[k] = init_state[size_t(i) * UPKIE_INIT_DIM + k];
  } else {
    const uint32_t ep = episode[i] + 1u;
    episode[i] = ep;
    sample_init_state(P, seed, env_offset + uint64_t(i), uint64_t(ep), init);
  }
  const BodyRecOut br{P.body_rec ? P.body_rec + i : nullptr, size_t(P.body_rec_stride)};
  if (P.spine_mode && lag) {
    SpineLag L;
    reset_robot_spine(P, S, L, init, eps, mu, WarpAny(), P.joint_limits, br);
    float lr[UPKIE_LAG_DIM];
    lag_to_row(L, lr);
    for (int k = 0; k < UPKIE_LAG_DIM; ++k) lag[size_t(k) * n_pad + i] = lr[k];
  } else {
    reset_robot(P, S, init, eps, mu, WarpAny(), P.joint_limits, br);
  }
  store_state(state, n_pad, i, S);
  err[i] = 0;
  done_prev[i] = 0;
}

__global__ void k_spine_obs(const __grid_constant__ SimParams P, int n, int n_pad, const float* __restrict__ state,
                            const uint32_t* __restrict__ tick, uint64_t env_offset, float* __restrict__ out,
                            const float* __restrict__ lag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float o[UPKIE_SPINE_DIM];
  const NoiseCtx nz{env_offset + uint64_t(i), tick[i]};  // same draw as the step that produced this state
  if (P.spine_mode && lag) {
    // the observation the spine assembled in the first cycle of the last step / the last cycle of the reset
    float lr[UPKIE_LAG_DIM];
    for (int k = 0; k < UPKIE_LAG_DIM; ++k) lr[k] = lag[size_t(k) * n_pad + i];
    SpineLag L;
    lag_from_row(lr, L);
    spine_observation_from_lag(P, L, o);
  } else {
    RobotState S;
    load_state(state, n_pad, i, S);
    float tq[6];
    measured_torques(P, S, &nz, tq);
    spine_observation(P, S, o, tq);
  }
  apply_imu_uncertainty(P, nz, o);
#pragma unroll
  for (int k = 0; k < UPKIE_SPINE_DIM; ++k) out[size_t(i) * UPKIE_SPINE_DIM + k] = o[k];
}

__global__ void k_reset_obs(const __grid_constant__ SimParams P, int n, int n_pad, const float* __restrict__ state,
                            const uint32_t* __restrict__ tick, uint64_t env_offset, int obs_dim,
                            float* __restrict__ out, const float* __restrict__ lag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RobotState S;
  load_state(state, n_pad, i, S);
  if (obs_dim == UPKIE_OBS_DIM && P.spine_mode && lag) {
    for (int j = 0; j < 6; ++j) {
      float* o = out + size_t(i) * UPKIE_OBS_DIM + j * 5;
      for (int k = 0; k < 3; ++k) o[k] = lag[size_t(UPKIE_LAG_OBS_REPLY + 3 * j + k) * n_pad + i];
      o[3] = 20.0f; o[4] = 18.0f;
    }
    return;
  }
  if (obs_dim == UPKIE_OBS_DIM) {
    float tq[6];
    const NoiseCtx nz{env_offset + uint64_t(i), tick[i]};
    measured_torques(P, S, &nz, tq);
    for (int j = 0; j < 6; ++j) {
      float* o = out + size_t(i) * UPKIE_OBS_DIM + j * 5;
      o[0] = S.q[j]; o[1] = S.qd[j]; o[2] = tq[j]; o[3] = 42.0f; o[4] = 18.0f;
    }
    return;
  }
  float o6[6];
  gyropod_obs(P, S, o6);
  if (obs_dim == 6) {
    for (int k = 0; k < 6; ++k) out[size_t(i) * 6 + k] = o6[k];
  } else {
    out[size_t(i) * 4 + 0] = o6[1]; out[size_t(i) * 4 + 1] = o6[0];
    out[size_t(i) * 4 + 2] = o6[4]; out[size_t(i) * 4 + 3] = o6[3];
  }
}

// SoA <-> AoS state transposes
__global__ void k_get_state(int n, int n_pad, const float* __restrict__ state, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < UPKIE_STATE_DIM; ++k) out[size_t(i) * UPKIE_STATE_DIM + k] = state[size_t(k) * n_pad + i];
}
__global__ void k_set_state(int n, int n_pad, float* __restrict__ state, const float* __restrict__ in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < UPKIE_STATE_DIM; ++k) state[size_t(k) * n_pad + i] = in[size_t(i) * UPKIE_STATE_DIM + k];
}
__global__ void k_set_ext(int n, int n_pad, float* __restrict__ ext, const float* __restrict__ in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < 3 * UPKIE_NB; ++k) ext[size_t(k) * n_pad + i] = in[size_t(i) * 3 * UPKIE_NB + k];
}
__global__ void k_init_state(const __grid_constant__ SimParams P, int n, int n_pad, float* __restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  for (int k = 0; k < UPKIE_STATE_DIM; ++k) state[size_t(k) * n_pad + i] = 0.f;
  for (int k = 0; k < 3; ++k) state[size_t(UPKIE_ST_POS + k) * n_pad + i] = P.init_pos[k];
  for (int k = 0; k < 4; ++k) state[size_t(UPKIE_ST_QUAT + k) * n_pad + i] = P.init_quat[k];
  (void)n;
}

// Block size per launch. Two measured effects (profiles/r01_variants.md): (1) the warps of a block run the
// substep body in lock-step (one barrier per substep) and share its instruction fetches, so large blocks win
// (256 > 128 > 64); (2) with 255 registers/thread one 256-thread block fills an SM, so 65 536 envs = 256 blocks
// = 1.73 waves left 25 % of the SM-time idle. Pick the number of waves k first, then the block size (multiple of
// 32) that makes the grid k * num_sms blocks: every SM gets the same number of equally sized blocks.
int pick_block(const Handle* h, int cnt) {
  if (h->block > 0) return h->block;
  const int per_wave = h->num_sms * UPKIE_MAX_THREADS;
  const int waves = (cnt + per_wave - 1) / per_wave;
  int block = (cnt + waves * h->num_sms - 1) / (waves * h->num_sms);
  block = (block + 31) / 32 * 32;
  if (block < 32) block = 32;
  if (block > UPKIE_MAX_THREADS) block = UPKIE_MAX_THREADS;
  return block;
}

// envs [i0, i0 + cnt): all buffers are indexed by the env index of the handle. `tile` selects the
// shared-memory-tile instantiation (host buffers), see kernel_common.cuh.
int step_range(Handle* h, int mode, int i0, int cnt, const float* action, float* obs, float* reward, uint8_t* term,
               uint8_t* trunc, cudaStream_t s, bool tile = false, bool persistent = true, bool compact = false,
               bool multicast = false, const PeerPtrs* peers = nullptr) {
  StepArgs a;
  std::memset(&a.peers, 0, sizeof(a.peers));
  if (peers) a.peers = *peers;
  a.P = &h->P;
  a.mode = mode;
  a.autoreset = h->autoreset;
  a.noise = h->P.joint_limits ? 2 : ((h->P.any_ctrl_noise || h->P.any_meas_noise || h->ext) ? 1 : 0);  // "extras" kernels
  if (h->P.body_contacts) {  // the model's collision points hold contact rows: the NOISE=4 kernels (step_*_body.cu)
    if (multicast) return fail(UPKIE_B200_EINVAL, "body_contacts has no in-kernel rollout transport (use upkie_b200_step_servos_compact)");
    a.noise = 4;
  }
  a.lag = nullptr;
  if (h->P.spine_mode) {
    if (mode != MODE_SERVOS) return fail(UPKIE_B200_EINVAL, "spine_mode supports UpkieServos steps only");
    if (multicast) return fail(UPKIE_B200_EINVAL, "spine_mode has no in-kernel rollout transport");
    a.noise = 3;
    a.lag = h->lag;
  }
  a.ext = h->ext;
  a.ext_local = h->ext_local;
  a.i0 = i0;
  a.cnt = cnt;
  a.n_pad = h->n_pad;
  a.block = tile ? h->host_block : pick_block(h, cnt);
  a.grid = (tile && persistent) ? h->num_sms * h->host_blocks_per_sm : 0;
  a.compact_obs = compact ? 1 : 0;
  a.state = h->state;
  a.action = action;
  a.obs = obs;
  a.reward = reward;
  a.terminated = term;
  a.truncated = trunc;
  a.eps = h->eps;
  a.mu = h->mu;
  a.err = h->err;
  a.done_prev = h->done_prev;
  a.episode = h->episode;
  a.tick = h->tick;
  a.seed = h->seed;
  a.env_offset = h->env_offset;
  a.stream = s;
  CUDA_TRY(multicast ? launch_step_multicast(a) : (tile ? launch_step_host(a) : launch_step_device(a)));
  h->step_launches += 1;
  return UPKIE_B200_OK;
}

int step_any(Handle* h, int mode, const float* action, float* obs, float* reward, uint8_t* term, uint8_t* trunc,
             cudaStream_t s) {
  if (!action || !obs || !term) return fail(UPKIE_B200_EINVAL, "step: null buffer");
  CUDA_TRY(cudaSetDevice(h->device));
  return step_range(h, mode, 0, h->n, action, obs, reward, term, trunc, s);
}

int ensure_staging(Handle* h) {
  if (h->d_act) return UPKIE_B200_OK;
  const size_t n = size_t(h->n);
  CUDA_TRY(cudaSetDevice(h->device));
  for (int k = 0; k < kHostStreams; ++k) CUDA_TRY(cudaStreamCreateWithFlags(&h->host_streams[k], cudaStreamNonBlocking));
  for (int k = 0; k < 64; ++k) CUDA_TRY(cudaEventCreateWithFlags(&h->host_events[k], cudaEventDisableTiming));
  CUDA_TRY(cudaMallocHost(&h->h_act, n * UPKIE_ACT_DIM * sizeof(float)));
  CUDA_TRY(cudaMallocHost(&h->h_obs, n * UPKIE_OBS_DIM * sizeof(float)));
  CUDA_TRY(cudaMallocHost(&h->h_rew, n * sizeof(float)));
  CUDA_TRY(cudaMallocHost(&h->h_term, n));
  CUDA_TRY(cudaMallocHost(&h->h_trunc, n));
  CUDA_TRY(cudaMalloc(&h->d_act, n * UPKIE_ACT_DIM * sizeof(float)));
  CUDA_TRY(cudaMalloc(&h->d_obs, n * UPKIE_OBS_DIM * sizeof(float)));
  CUDA_TRY(cudaMalloc(&h->d_rew, n * sizeof(float)));
  CUDA_TRY(cudaMalloc(&h->d_term, n));
  CUDA_TRY(cudaMalloc(&h->d_trunc, n));
  return UPKIE_B200_OK;
}

// device-side alias of a pinned host buffer (nullptr when the buffer is pageable or not mapped)
template <typename T>
T* mapped(T* p) {
  if (!p) return nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return static_cast<T*>(at.devicePointer);
}

// Host-buffer step. Pageable caller buffers are first staged through the handle's pinned buffers; on
// pinned (mapped) buffers one of three pipelines runs, `zero_copy` selecting it:
//   2 (default, servos)  hybrid: the copy engine streams the action rows in, chunk by chunk on one stream; each
//                        chunk's TILE=1 kernel waits for its rows only and writes observations / flags straight
//                        to host memory. Copy-engine reads overlap SM writes on the link, SM reads do not
//                        (63 GB/s combined, tools/micro/pcie_duplex.cu).
//   1                    one persistent TILE=1 launch reading actions from and writing observations to host memory.
//   0                    H2D copy -> TILE=0 kernel -> D2H copies per chunk on rotating streams.
// `compact` (servos): observation rows [6][3] = position, velocity, torque (TILE=1 kernels).
int step_host(Handle* h, int mode, const float* action, float* obs, float* reward, uint8_t* term, uint8_t* trunc,
              bool compact = false) {
  if (!action || !obs || !term) return fail(UPKIE_B200_EINVAL, "step_host: null buffer");
  int rc = ensure_staging(h);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t n = size_t(h->n);
  const size_t act_dim = mode == MODE_SERVOS ? UPKIE_ACT_DIM : (mode == MODE_GYROPOD ? 2 : 1);
  const size_t obs_dim = mode == MODE_SERVOS ? (compact ? 18 : UPKIE_OBS_DIM) : (mode == MODE_GYROPOD ? 6 : 4);
  // reward / truncated are constants of the reference (0.0 and false): callers may pass NULL for them
  const bool pin_in = mapped(action) != nullptr;
  const bool pin_out = mapped(obs) && (!reward || mapped(reward)) && mapped(term) && (!trunc || mapped(trunc));
  if (!pin_in) std::memcpy(h->h_act, action, n * act_dim * sizeof(float));
  const float* src_act = pin_in ? action : h->h_act;
  float* dst_obs = pin_out ? obs : h->h_obs;
  float* dst_rew = reward ? (pin_out ? reward : h->h_rew) : nullptr;
  uint8_t* dst_term = pin_out ? term : h->h_term;
  uint8_t* dst_trunc = trunc ? (pin_out ? trunc : h->h_trunc) : nullptr;

  // Order the private (non-blocking) streams after whatever the caller enqueued on the default stream - reset(),
  // set_state(), set_counters() through PyTorch's default current stream: without this a step_*_host() right after
  // an asynchronous reset raced with it (round-1 advisor finding). Callers on other streams synchronise themselves.
  CUDA_TRY(cudaEventRecord(h->host_events[63], cudaStreamLegacy));
  for (int k = 0; k < kHostStreams; ++k) CUDA_TRY(cudaStreamWaitEvent(h->host_streams[k], h->host_events[63], 0));

  int pipeline = h->zero_copy;
  if (pipeline == 2 && mode != MODE_SERVOS) pipeline = 1;  // tiny rows: nothing to stream
  // chunk boundaries (multiples of 256 envs). Default: host_chunks equal chunks; UPKIE_B200_HOST_SPLIT gives the
  // fractions explicitly (a small last chunk shortens the exposed tail: last kernel + its write drain).
  int start[65];
  int chunks = 0;
  start[0] = 0;
  if (h->host_split_n > 0 && h->n >= 4 * 8192) {
    double acc = 0.0;
    for (int c = 0; c < h->host_split_n && chunks < 62; ++c) {
      acc += h->host_split[c];
      int end = c + 1 == h->host_split_n ? h->n : int(acc * h->n + 0.5);
      end = (end + 255) / 256 * 256;
      if (end > h->n) end = h->n;
      if (end > start[chunks]) start[++chunks] = end;
    }
    if (start[chunks] < h->n) start[++chunks] = h->n;
  } else {
    const int want = h->n >= 4 * 8192 ? h->host_chunks : (h->n >= 2 * 8192 ? 2 : 1);
    int per = (h->n + want - 1) / want;
    per = (per + 255) / 256 * 256;
    for (int i0 = 0; i0 < h->n && chunks < 63; i0 += per) start[++chunks] = (i0 + per < h->n) ? i0 + per : h->n;
  }
  auto chunk_count = [&](int c) { return start[c + 1] - start[c]; };

  if (pipeline == 2) {
    cudaStream_t sc = h->host_streams[0];
    for (int c = 0; c < chunks; ++c) {
      const size_t i0 = size_t(start[c]);
      CUDA_TRY(cudaMemcpyAsync(h->d_act + i0 * act_dim, src_act + i0 * act_dim, size_t(chunk_count(c)) * act_dim * sizeof(float),
                               cudaMemcpyHostToDevice, sc));
      CUDA_TRY(cudaEventRecord(h->host_events[c], sc));
    }
    for (int c = 0; c < chunks; ++c) {
      cudaStream_t sk = h->host_streams[1 + (c % h->host_kernel_streams)];
      CUDA_TRY(cudaStreamWaitEvent(sk, h->host_events[c], 0));
      rc = step_range(h, mode, start[c], chunk_count(c), h->d_act, mapped(dst_obs), mapped(dst_rew), mapped(dst_term),
                      mapped(dst_trunc), sk, /*tile=*/true, /*persistent=*/false, compact);
      if (rc) return rc;
    }
    for (int k = 0; k < h->host_kernel_streams; ++k) CUDA_TRY(cudaStreamSynchronize(h->host_streams[1 + k]));
  } else if (pipeline == 1) {
    cudaStream_t s = h->host_streams[0];
    rc = step_range(h, mode, 0, h->n, mapped(src_act), mapped(dst_obs), mapped(dst_rew), mapped(dst_term),
                    mapped(dst_trunc), s, /*tile=*/true, /*persistent=*/true, compact);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(s));
  } else {
    for (int c = 0; c < chunks; ++c) {
      const int i0 = start[c];
      const int cnt = chunk_count(c);
      cudaStream_t s = h->host_streams[c % kHostStreams];
      CUDA_TRY(cudaMemcpyAsync(h->d_act + size_t(i0) * act_dim, src_act + size_t(i0) * act_dim,
                               size_t(cnt) * act_dim * sizeof(float), cudaMemcpyHostToDevice, s));
      // compact rows exist in the TILE=1 kernels only; device staging buffers either way
      rc = step_range(h, mode, i0, cnt, h->d_act, h->d_obs, h->d_rew, h->d_term, h->d_trunc, s, /*tile=*/compact,
                      /*persistent=*/false, compact);
      if (rc) return rc;
      CUDA_TRY(cudaMemcpyAsync(dst_obs + size_t(i0) * obs_dim, h->d_obs + size_t(i0) * obs_dim,
                               size_t(cnt) * obs_dim * sizeof(float), cudaMemcpyDeviceToHost, s));
      if (dst_rew) CUDA_TRY(cudaMemcpyAsync(dst_rew + i0, h->d_rew + i0, size_t(cnt) * sizeof(float), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaMemcpyAsync(dst_term + i0, h->d_term + i0, size_t(cnt), cudaMemcpyDeviceToHost, s));
      if (dst_trunc) CUDA_TRY(cudaMemcpyAsync(dst_trunc + i0, h->d_trunc + i0, size_t(cnt), cudaMemcpyDeviceToHost, s));
    }
    for (int k = 0; k < kHostStreams; ++k) CUDA_TRY(cudaStreamSynchronize(h->host_streams[k]));
  }
  if (!pin_out) {
    std::memcpy(obs, h->h_obs, n * obs_dim * sizeof(float));
    if (reward) std::memcpy(reward, h->h_rew, n * sizeof(float));
    std::memcpy(term, h->h_term, n);
    if (trunc) std::memcpy(trunc, h->h_trunc, n);
  }
  return UPKIE_B200_OK;
}

}  // namespace

// ---- C ABI ----------------------------------------------------------------------------

extern "C" {

int upkie_b200_abi_version(void) { return UPKIE_B200_ABI_VERSION; }
const char* upkie_b200_last_error(void) { return g_last_error.c_str(); }

int upkie_b200_default_config(UpkieSimConfig* config) {
  if (!config) return fail(UPKIE_B200_EINVAL, "default_config: null");
  default_sim_config(config);
  return UPKIE_B200_OK;
}
int upkie_b200_default_mpc_config(UpkieMpcConfig* config) {
  if (!config) return fail(UPKIE_B200_EINVAL, "default_mpc_config: null");
  default_mpc_config(config);
  return UPKIE_B200_OK;
}

int upkie_b200_create(const UpkieModel* model, const UpkieSimConfig* config, int n_envs, int device, void** handle) {
  if (!model || !config || !handle) return fail(UPKIE_B200_EINVAL, "create: null argument");
  if (n_envs < 1) return fail(UPKIE_B200_EINVAL, "create: n_envs must be >= 1");
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(UPKIE_B200_ECUDA, "create: no CUDA device available (this library has no CPU path)");
  if (device < 0 || device >= count) return fail(UPKIE_B200_EINVAL, "create: invalid device index");
  Handle* h = new (std::nothrow) Handle();
  if (!h) return fail(UPKIE_B200_ENOMEM, "create: out of host memory");
  std::memset(&h->P, 0, sizeof(h->P));
  std::string err;
  int rc = make_sim_params(*model, *config, h->P, err);
  if (rc) { delete h; return fail(rc, err); }
  h->magic = kMagic;
  h->model = *model;
  h->n = n_envs;
  h->n_pad = (n_envs + 31) / 32 * 32;
  h->device = device;
  if (const char* b = std::getenv("UPKIE_B200_BLOCK")) {
    const int v = std::atoi(b);
    if (v >= 32 && v <= UPKIE_MAX_THREADS && v % 32 == 0) h->block = v;
  }
  if (const char* b = std::getenv("UPKIE_B200_HOST_BLOCK")) {  // developer knob
    const int v = std::atoi(b);
    if (v >= 32 && v <= 160 && v % 32 == 0) h->host_block = v;  // 2 x 4608 B of tile per warp: <= 48 KB
  }
  if (const char* b = std::getenv("UPKIE_B200_HOST_BLOCKS_PER_SM")) {  // developer knob
    const int v = std::atoi(b);
    if (v >= 1 && v <= 8) h->host_blocks_per_sm = v;
  }
  if (const char* b = std::getenv("UPKIE_B200_ZERO_COPY")) h->zero_copy = std::atoi(b);  // developer knob: 0, 1, 2
  if (const char* b = std::getenv("UPKIE_B200_HOST_SPLIT")) {  // developer knob: "0.4,0.4,0.2"
    h->host_split_n = 0;
    const char* p = b;
    while (*p && h->host_split_n < 8) {
      char* end = nullptr;
      const double v = std::strtod(p, &end);
      if (end == p) break;
      if (v > 0.0) h->host_split[h->host_split_n++] = v;
      p = (*end == ',') ? end + 1 : end;
    }
  }
  if (const char* b = std::getenv("UPKIE_B200_HOST_KERNEL_STREAMS")) {  // developer knob
    const int v = std::atoi(b);
    if (v >= 1 && v <= 2) h->host_kernel_streams = v;
  }
  if (const char* b = std::getenv("UPKIE_B200_HOST_CHUNKS")) {  // developer knob
    const int v = std::atoi(b);
    if (v >= 1 && v <= 62) h->host_chunks = v;
  }
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess) e = cudaMalloc(&h->state, size_t(UPKIE_STATE_DIM) * h->n_pad * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->err, size_t(n_envs) * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&h->done_prev, size_t(n_envs));
  if (e == cudaSuccess) e = cudaMalloc(&h->episode, size_t(n_envs) * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(h->err, 0, size_t(n_envs) * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(h->done_prev, 0, size_t(n_envs));
  if (e == cudaSuccess) e = cudaMemset(h->episode, 0, size_t(n_envs) * sizeof(uint32_t));
  if (e == cudaSuccess && h->P.spine_mode) {
    e = cudaMalloc(&h->lag, size_t(UPKIE_LAG_DIM) * h->n_pad * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(h->lag, 0, size_t(UPKIE_LAG_DIM) * h->n_pad * sizeof(float));
  }
  if (e == cudaSuccess && h->P.body_contacts) {
    e = cudaMalloc(&h->body_rec, size_t(UPKIE_BODY_REC_DIM) * h->n_pad * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(h->body_rec, 0, size_t(UPKIE_BODY_REC_DIM) * h->n_pad * sizeof(float));
    h->P.body_rec = h->body_rec;
    h->P.body_rec_stride = h->n_pad;
  }
  if (e == cudaSuccess) e = cudaMalloc(&h->tick, size_t(n_envs) * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(h->tick, 0, size_t(n_envs) * sizeof(uint32_t));
  if (e == cudaSuccess) {
    k_init_state<<<(h->n_pad + 127) / 128, 128>>>(h->P, h->n, h->n_pad, h->state);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    std::string msg = std::string("create: ") + cudaGetErrorString(e);
    upkie_b200_destroy(h);
    return fail(UPKIE_B200_ECUDA, msg);
  }
  *handle = h;
  return UPKIE_B200_OK;
}

void upkie_b200_destroy(void* handle) {
  Handle* h = as_handle(handle);
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->state); cudaFree(h->eps); cudaFree(h->mu); cudaFree(h->err); cudaFree(h->done_prev); cudaFree(h->episode);
  cudaFree(h->tick); cudaFree(h->ext); cudaFree(h->lag); cudaFree(h->body_rec);
  cudaFreeHost(h->h_act); cudaFreeHost(h->h_obs); cudaFreeHost(h->h_rew); cudaFreeHost(h->h_term); cudaFreeHost(h->h_trunc);
  cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_rew); cudaFree(h->d_term); cudaFree(h->d_trunc);
  for (int k = 0; k < kHostStreams; ++k)
    if (h->host_streams[k]) cudaStreamDestroy(h->host_streams[k]);
  for (int k = 0; k < 64; ++k)
    if (h->host_events[k]) cudaEventDestroy(h->host_events[k]);
  h->magic = 0;
  delete h;
}

int upkie_b200_num_envs(void* handle) {
  Handle* h = as_handle(handle);
  return h ? h->n : fail(UPKIE_B200_EINVAL, "invalid handle");
}

int upkie_b200_set_config(void* handle, const UpkieSimConfig* config) {
  Handle* h = as_handle(handle);
  if (!h || !config) return fail(UPKIE_B200_EINVAL, "set_config: invalid argument");
  SimParams P;
  std::memset(&P, 0, sizeof(P));
  std::string err;
  int rc = make_sim_params(h->model, *config, P, err);
  if (rc) return fail(rc, err);
  if (P.spine_mode != h->P.spine_mode) return fail(UPKIE_B200_EINVAL, "set_config: spine_mode is fixed at creation");
  if (P.body_contacts && !h->body_rec) {  // switched on after creation: the record buffer is allocated now
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMalloc(&h->body_rec, size_t(UPKIE_BODY_REC_DIM) * h->n_pad * sizeof(float)));
    CUDA_TRY(cudaMemset(h->body_rec, 0, size_t(UPKIE_BODY_REC_DIM) * h->n_pad * sizeof(float)));
  }
  P.body_rec = P.body_contacts ? h->body_rec : nullptr;
  P.body_rec_stride = h->n_pad;
  // kernels read the parameter block by value at launch: steps already enqueued keep the old one
  h->P = P;
  return UPKIE_B200_OK;
}

int upkie_b200_set_autoreset(void* handle, int mode, uint64_t seed, uint64_t env_offset) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  if (mode < 0 || mode > 2) return fail(UPKIE_B200_EINVAL, "set_autoreset: mode must be 0 (disabled), 1 (next step) or 2 (same step)");
  h->autoreset = mode;
  h->seed = seed;
  h->env_offset = env_offset;
  return UPKIE_B200_OK;
}

int upkie_b200_set_randomization(void* handle, const float* friction, const float* inertia_eps, void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaSetDevice(h->device));
  if (friction) {
    if (!h->mu) CUDA_TRY(cudaMalloc(&h->mu, size_t(h->n) * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(h->mu, friction, size_t(h->n) * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else if (h->mu) {
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaFree(h->mu);
    h->mu = nullptr;
  }
  if (inertia_eps) {
    if (!h->eps) CUDA_TRY(cudaMalloc(&h->eps, size_t(h->n) * 6 * sizeof(float)));
    CUDA_TRY(cudaMemcpyAsync(h->eps, inertia_eps, size_t(h->n) * 6 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else if (h->eps) {
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaFree(h->eps);
    h->eps = nullptr;
  }
  return UPKIE_B200_OK;
}

int upkie_b200_reset(void* handle, const uint8_t* mask, const float* init_state, uint64_t seed, uint64_t env_offset,
                     void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rblock = 128;
  const int grid = (h->n + rblock - 1) / rblock;
  k_reset<<<grid, rblock, 0, s>>>(h->P, h->n, h->n_pad, h->state, mask, init_state, h->eps, h->mu, h->err,
                                    h->done_prev, h->episode, seed, env_offset, h->lag);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}

int upkie_b200_step_servos(void* handle, const float* action, float* obs, float* reward, uint8_t* terminated,
                           uint8_t* truncated, void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  return step_any(h, MODE_SERVOS, action, obs, reward, terminated, truncated, static_cast<cudaStream_t>(stream));
}

int upkie_b200_step_servos_compact(void* handle, const float* action, float* obs, uint8_t* terminated, void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  if (!action || !obs || !terminated) return fail(UPKIE_B200_EINVAL, "step_servos_compact: null buffer");
  CUDA_TRY(cudaSetDevice(h->device));
  return step_range(h, MODE_SERVOS, 0, h->n, action, obs, nullptr, terminated, nullptr, static_cast<cudaStream_t>(stream),
                    /*tile=*/true, /*persistent=*/false, /*compact=*/true);
}

int upkie_b200_step_servos_multicast(void* handle, const float* action, float* obs_mc, uint8_t* terminated_mc, void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  if (!action || !obs_mc || !terminated_mc) return fail(UPKIE_B200_EINVAL, "step_servos_multicast: null buffer");
  if (h->n % 32 != 0) return fail(UPKIE_B200_EINVAL, "step_servos_multicast: the number of envs must be a multiple of 32");
  if (((reinterpret_cast<uintptr_t>(action) | reinterpret_cast<uintptr_t>(obs_mc)) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(terminated_mc) & 3) != 0)
    return fail(UPKIE_B200_EINVAL, "step_servos_multicast: action / obs rows must be 16-byte, terminated 4-byte aligned");
  CUDA_TRY(cudaSetDevice(h->device));
  return step_range(h, MODE_SERVOS, 0, h->n, action, obs_mc, nullptr, terminated_mc, nullptr, static_cast<cudaStream_t>(stream),
                    /*tile=*/true, /*persistent=*/false, /*compact=*/true, /*multicast=*/true);
}

int upkie_b200_step_servos_peers(void* handle, const float* action, float* const* obs_ptrs,
                                 uint8_t* const* terminated_ptrs, int n_peers, void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  if (!action || !obs_ptrs || !terminated_ptrs) return fail(UPKIE_B200_EINVAL, "step_servos_peers: null buffer");
  if (n_peers < 1 || n_peers > UPKIE_MAX_PEERS) return fail(UPKIE_B200_EINVAL, "step_servos_peers: 1 <= n_peers <= UPKIE_MAX_PEERS");
  if (h->n % 32 != 0) return fail(UPKIE_B200_EINVAL, "step_servos_peers: the number of envs must be a multiple of 32");
  PeerPtrs pp;
  std::memset(&pp, 0, sizeof(pp));
  pp.n = n_peers;
  for (int p = 0; p < UPKIE_MAX_PEERS; ++p) {
    pp.obs[p] = p < n_peers ? obs_ptrs[p] : nullptr;
    pp.term[p] = p < n_peers ? terminated_ptrs[p] : nullptr;
    if (p < n_peers && (!pp.obs[p] || !pp.term[p] || (reinterpret_cast<uintptr_t>(pp.obs[p]) & 15) != 0 ||
                        (reinterpret_cast<uintptr_t>(pp.term[p]) & 3) != 0))
      return fail(UPKIE_B200_EINVAL, "step_servos_peers: slot pointers must be non-null, obs 16-byte and terminated 4-byte aligned");
  }
  if ((reinterpret_cast<uintptr_t>(action) & 15) != 0) return fail(UPKIE_B200_EINVAL, "step_servos_peers: action rows must be 16-byte aligned");
  CUDA_TRY(cudaSetDevice(h->device));
  return step_range(h, MODE_SERVOS, 0, h->n, action, pp.obs[0], nullptr, pp.term[0], nullptr, static_cast<cudaStream_t>(stream),
                    /*tile=*/true, /*persistent=*/false, /*compact=*/true, /*multicast=*/true, &pp);
}

namespace {
int fill_push(const Handle* h, const UpkiePush* push, PeerPtrs& pp, const char* who) {
  std::memset(&pp, 0, sizeof(pp));
  pp.deferred = 1;
  if (!push) return UPKIE_B200_OK;  // nothing to send
  if (push->n_peers < 0 || push->n_peers > UPKIE_MAX_PEERS) return fail(UPKIE_B200_EINVAL, std::string(who) + ": 0 <= n_peers <= UPKIE_MAX_PEERS");
  if (h->n % 32 != 0) return fail(UPKIE_B200_EINVAL, std::string(who) + ": the number of envs must be a multiple of 32");
  auto bad = [](const void* q, uintptr_t mask) { return !q || (reinterpret_cast<uintptr_t>(q) & mask) != 0; };
  if (push->src_obs) {
    if (bad(push->src_obs, 15) || bad(push->src_terminated, 3)) return fail(UPKIE_B200_EINVAL, std::string(who) + ": source slot misaligned or null");
    pp.src_obs = push->src_obs;
    pp.src_term = push->src_terminated;
    pp.n = push->n_peers;
    if (pp.n == 0) {
      if (bad(push->mc_obs, 15) || bad(push->mc_terminated, 3)) return fail(UPKIE_B200_EINVAL, std::string(who) + ": multicast slot misaligned or null");
      pp.mc_obs = push->mc_obs;
      pp.mc_term = push->mc_terminated;
    }
    for (int p = 0; p < pp.n; ++p) {
      if (bad(push->peer_obs[p], 15) || bad(push->peer_terminated[p], 3)) return fail(UPKIE_B200_EINVAL, std::string(who) + ": peer slot misaligned or null");
      pp.obs[p] = push->peer_obs[p];
      pp.term[p] = push->peer_terminated[p];
    }
  }
  return UPKIE_B200_OK;
}
}  // namespace

int upkie_b200_step_servos_push(void* handle, const float* action, float* obs, uint8_t* terminated,
                                const UpkiePush* push, void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  if (!action || !obs || !terminated) return fail(UPKIE_B200_EINVAL, "step_servos_push: null buffer");
  if (h->n % 32 != 0) return fail(UPKIE_B200_EINVAL, "step_servos_push: the number of envs must be a multiple of 32");
  if (((reinterpret_cast<uintptr_t>(action) | reinterpret_cast<uintptr_t>(obs)) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(terminated) & 3) != 0)
    return fail(UPKIE_B200_EINVAL, "step_servos_push: action / obs rows must be 16-byte, terminated 4-byte aligned");
  PeerPtrs pp;
  int rc = fill_push(h, push, pp, "step_servos_push");
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  return step_range(h, MODE_SERVOS, 0, h->n, action, obs, nullptr, terminated, nullptr, static_cast<cudaStream_t>(stream),
                    /*tile=*/true, /*persistent=*/false, /*compact=*/true, /*multicast=*/true, &pp);
}

int upkie_b200_push_rows(void* handle, const UpkiePush* push, void* stream) {
  Handle* h = as_handle(handle);
  if (!h || !push || !push->src_obs) return fail(UPKIE_B200_EINVAL, "push_rows: invalid argument");
  PeerPtrs pp;
  int rc = fill_push(h, push, pp, "push_rows");
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(launch_push_rows(pp, h->n, static_cast<cudaStream_t>(stream)));
  return UPKIE_B200_OK;
}

int upkie_b200_step_gyropod(void* handle, const float* action, int act_dim, float* obs, float* reward,
                            uint8_t* terminated, uint8_t* truncated, void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  if (act_dim != 1 && act_dim != 2) return fail(UPKIE_B200_EINVAL, "step_gyropod: act_dim must be 1 (pendulum) or 2 (gyropod)");
  return step_any(h, act_dim == 2 ? MODE_GYROPOD : MODE_PENDULUM, action, obs, reward, terminated, truncated,
                  static_cast<cudaStream_t>(stream));
}

int upkie_b200_step_servos_host(void* handle, const float* action, float* obs, float* reward, uint8_t* terminated,
                                uint8_t* truncated) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  return step_host(h, MODE_SERVOS, action, obs, reward, terminated, truncated);
}
int upkie_b200_step_servos_host_compact(void* handle, const float* action, float* obs, uint8_t* terminated) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "step_servos_host_compact: bad handle");
  return step_host(h, MODE_SERVOS, action, obs, nullptr, terminated, nullptr, /*compact=*/true);
}

int upkie_b200_step_gyropod_host(void* handle, const float* action, int act_dim, float* obs, float* reward,
                                 uint8_t* terminated, uint8_t* truncated) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "invalid handle");
  if (act_dim != 1 && act_dim != 2) return fail(UPKIE_B200_EINVAL, "step_gyropod_host: act_dim must be 1 or 2");
  return step_host(h, act_dim == 2 ? MODE_GYROPOD : MODE_PENDULUM, action, obs, reward, terminated, truncated);
}

int upkie_b200_spine_obs(void* handle, float* out, void* stream) {
  Handle* h = as_handle(handle);
  if (!h || !out) return fail(UPKIE_B200_EINVAL, "spine_obs: invalid argument");
  CUDA_TRY(cudaSetDevice(h->device));
  k_spine_obs<<<(h->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->P, h->n, h->n_pad, h->state, h->tick, h->env_offset, out, h->lag);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}

int upkie_b200_reset_obs(void* handle, int obs_dim, float* obs, void* stream) {
  Handle* h = as_handle(handle);
  if (!h || !obs) return fail(UPKIE_B200_EINVAL, "reset_obs: invalid argument");
  if (obs_dim != 4 && obs_dim != 6 && obs_dim != UPKIE_OBS_DIM) return fail(UPKIE_B200_EINVAL, "reset_obs: obs_dim must be 4, 6 or 30");
  CUDA_TRY(cudaSetDevice(h->device));
  k_reset_obs<<<(h->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->P, h->n, h->n_pad, h->state, h->tick, h->env_offset, obs_dim, obs, h->lag);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}

int upkie_b200_get_state(void* handle, float* state, void* stream) {
  Handle* h = as_handle(handle);
  if (!h || !state) return fail(UPKIE_B200_EINVAL, "get_state: invalid argument");
  CUDA_TRY(cudaSetDevice(h->device));
  k_get_state<<<(h->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->n, h->n_pad, h->state, state);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}

int upkie_b200_set_state(void* handle, const float* state, void* stream) {
  Handle* h = as_handle(handle);
  if (!h || !state) return fail(UPKIE_B200_EINVAL, "set_state: invalid argument");
  CUDA_TRY(cudaSetDevice(h->device));
  k_set_state<<<(h->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->n, h->n_pad, h->state, state);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}

__global__ void k_lag_copy(int n, int n_pad, float* __restrict__ lag, float* __restrict__ rows, int to_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < UPKIE_LAG_DIM; ++k) {
    if (to_rows) rows[size_t(i) * UPKIE_LAG_DIM + k] = lag[size_t(k) * n_pad + i];
    else lag[size_t(k) * n_pad + i] = rows[size_t(i) * UPKIE_LAG_DIM + k];
  }
}

__global__ void k_body_rec_rows(int n, int n_pad, const float* __restrict__ rec, float* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < UPKIE_BODY_REC_DIM; ++k) rows[size_t(i) * UPKIE_BODY_REC_DIM + k] = rec ? rec[size_t(k) * n_pad + i] : 0.f;
}

int upkie_b200_get_body_contacts(void* handle, float* rows, void* stream) {
  Handle* h = as_handle(handle);
  if (!h || !rows) return fail(UPKIE_B200_EINVAL, "get_body_contacts: invalid argument");
  CUDA_TRY(cudaSetDevice(h->device));
  k_body_rec_rows<<<(h->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      h->n, h->n_pad, h->P.body_contacts ? h->body_rec : nullptr, rows);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}

int upkie_b200_get_lag(void* handle, float* lag_rows, void* stream) {
  Handle* h = as_handle(handle);
  if (!h || !lag_rows) return fail(UPKIE_B200_EINVAL, "get_lag: invalid argument");
  if (!h->lag) return fail(UPKIE_B200_EINVAL, "get_lag: the handle was not created with spine_mode");
  CUDA_TRY(cudaSetDevice(h->device));
  k_lag_copy<<<(h->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->n, h->n_pad, h->lag, lag_rows, 1);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}
int upkie_b200_set_lag(void* handle, const float* lag_rows, void* stream) {
  Handle* h = as_handle(handle);
  if (!h || !lag_rows) return fail(UPKIE_B200_EINVAL, "set_lag: invalid argument");
  if (!h->lag) return fail(UPKIE_B200_EINVAL, "set_lag: the handle was not created with spine_mode");
  CUDA_TRY(cudaSetDevice(h->device));
  k_lag_copy<<<(h->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->n, h->n_pad, h->lag, const_cast<float*>(lag_rows), 0);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}

int upkie_b200_set_external_forces(void* handle, const float* force, uint32_t local_mask, void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "set_external_forces: invalid handle");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  h->ext_local = local_mask;
  if (!force) {
    if (h->ext) {
      CUDA_TRY(cudaStreamSynchronize(s));
      cudaFree(h->ext);
      h->ext = nullptr;
    }
    return UPKIE_B200_OK;
  }
  if (!h->ext) {
    CUDA_TRY(cudaMalloc(&h->ext, size_t(3 * UPKIE_NB) * h->n_pad * sizeof(float)));
    CUDA_TRY(cudaMemsetAsync(h->ext, 0, size_t(3 * UPKIE_NB) * h->n_pad * sizeof(float), s));
  }
  k_set_ext<<<(h->n + 127) / 128, 128, 0, s>>>(h->n, h->n_pad, h->ext, force);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}

int upkie_b200_launch_count(void* handle, uint64_t* count) {
  Handle* h = as_handle(handle);
  if (!h || !count) return fail(UPKIE_B200_EINVAL, "launch_count: bad argument");
  *count = h->step_launches;
  return UPKIE_B200_OK;
}

int upkie_b200_error_flags(void* handle, uint32_t* flags, void* stream) {
  Handle* h = as_handle(handle);
  if (!h || !flags) return fail(UPKIE_B200_EINVAL, "error_flags: invalid argument");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(flags, h->err, size_t(h->n) * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                           static_cast<cudaStream_t>(stream)));
  return UPKIE_B200_OK;
}

int upkie_b200_get_counters(void* handle, uint32_t* episode, uint32_t* tick, uint8_t* pending_reset,
                            uint32_t* error_flags, void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "get_counters: invalid handle");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = size_t(h->n);
  if (episode) CUDA_TRY(cudaMemcpyAsync(episode, h->episode, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  if (tick) CUDA_TRY(cudaMemcpyAsync(tick, h->tick, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  if (pending_reset) CUDA_TRY(cudaMemcpyAsync(pending_reset, h->done_prev, n, cudaMemcpyDeviceToDevice, s));
  if (error_flags) CUDA_TRY(cudaMemcpyAsync(error_flags, h->err, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  return UPKIE_B200_OK;
}

int upkie_b200_set_counters(void* handle, const uint32_t* episode, const uint32_t* tick, const uint8_t* pending_reset,
                            const uint32_t* error_flags, void* stream) {
  Handle* h = as_handle(handle);
  if (!h) return fail(UPKIE_B200_EINVAL, "set_counters: invalid handle");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = size_t(h->n);
  if (episode) CUDA_TRY(cudaMemcpyAsync(h->episode, episode, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  if (tick) CUDA_TRY(cudaMemcpyAsync(h->tick, tick, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  if (pending_reset) CUDA_TRY(cudaMemcpyAsync(h->done_prev, pending_reset, n, cudaMemcpyDeviceToDevice, s));
  if (error_flags) CUDA_TRY(cudaMemcpyAsync(h->err, error_flags, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  return UPKIE_B200_OK;
}

// ---- MPC ---------------------------------------------------------------------------------

int upkie_b200_mpc_create(const UpkieMpcConfig* config, int n_robots, int device, void** mpc) {
  if (!config || !mpc) return fail(UPKIE_B200_EINVAL, "mpc_create: null argument");
  std::string err;
  int rc = mpc_create_impl(*config, n_robots, device, mpc, err);
  return rc ? fail(rc, err) : UPKIE_B200_OK;
}
void upkie_b200_mpc_destroy(void* mpc) { mpc_destroy_impl(mpc); }
int upkie_b200_mpc_reset(void* mpc, const uint8_t* mask, void* stream) {
  std::string err;
  int rc = mpc_reset_impl(mpc, mask, static_cast<cudaStream_t>(stream), err);
  return rc ? fail(rc, err) : UPKIE_B200_OK;
}
int upkie_b200_mpc_step(void* mpc, const float* x0, const float* v_target, const uint8_t* floor_contact, float dt,
                        float* v_cmd, float* first_input, uint8_t* found, void* stream) {
  std::string err;
  int rc = mpc_step_impl(mpc, x0, v_target, floor_contact, dt, v_cmd, first_input, found,
                         static_cast<cudaStream_t>(stream), err);
  return rc ? fail(rc, err) : UPKIE_B200_OK;
}
int upkie_b200_mpc_plan(void* mpc, float* plan, void* stream) {
  std::string err;
  int rc = mpc_plan_impl(mpc, plan, static_cast<cudaStream_t>(stream), err);
  return rc ? fail(rc, err) : UPKIE_B200_OK;
}

// ---- observer pipeline --------------------------------------------------------------------

int upkie_b200_default_observer_config(const UpkieModel* model, UpkieObserverConfig* config) {
  if (!model || !config) return fail(UPKIE_B200_EINVAL, "default_observer_config: null");
  default_observer_config(*model, config);
  return UPKIE_B200_OK;
}
int upkie_b200_default_wheel_balancer_config(UpkieWheelBalancerConfig* config) {
  if (!config) return fail(UPKIE_B200_EINVAL, "default_wheel_balancer_config: null");
  default_wheel_balancer_config(config);
  return UPKIE_B200_OK;
}
int upkie_b200_wheel_balancer_create(const UpkieWheelBalancerConfig* config, int n_robots, int device, void** balancer) {
  if (!config || !balancer) return fail(UPKIE_B200_EINVAL, "wheel_balancer_create: null argument");
  std::string err;
  int rc = wheel_balancer_create_impl(*config, n_robots, device, balancer, err);
  return rc ? fail(rc, err) : UPKIE_B200_OK;
}
void upkie_b200_wheel_balancer_destroy(void* balancer) { wheel_balancer_destroy_impl(balancer); }
int upkie_b200_wheel_balancer_reset(void* balancer, const uint8_t* mask, void* stream) {
  WheelBalancerHandle* h = as_wheel_balancer(balancer);
  if (!h) return fail(UPKIE_B200_EINVAL, "wheel_balancer_reset: invalid handle");
  CUDA_TRY(cudaSetDevice(h->device));
  k_wheel_balancer_reset<<<(h->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(h->n, mask, h->state);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}
int upkie_b200_wheel_balancer_step(void* balancer, const float* obs, int obs_layout, const float* target, float* action,
                                   void* stream) {
  WheelBalancerHandle* h = as_wheel_balancer(balancer);
  if (!h || !obs || !action) return fail(UPKIE_B200_EINVAL, "wheel_balancer_step: invalid argument");
  int stride, pitch, contact, odom;
  if (obs_layout == UPKIE_OBS_LAYOUT_SPINE) {
    stride = UPKIE_SPINE_DIM; pitch = UPKIE_SP_PITCH; contact = UPKIE_SP_CONTACT; odom = UPKIE_SP_ODOM_POS;
  } else if (obs_layout == UPKIE_OBS_LAYOUT_OBSERVERS) {
    stride = UPKIE_OBSV_DIM; pitch = UPKIE_OBSV_PITCH; contact = UPKIE_OBSV_CONTACT; odom = UPKIE_OBSV_ODOM_POS;
  } else {
    return fail(UPKIE_B200_EINVAL, "wheel_balancer_step: unknown observation layout");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  k_wheel_balancer_step<<<(h->n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      h->P, h->n, h->state, obs, stride, pitch, contact, odom, target, action);
  CUDA_TRY(cudaGetLastError());
  return UPKIE_B200_OK;
}
int upkie_b200_wheel_balancer_state(void* balancer, float* state, void* stream) {
  WheelBalancerHandle* h = as_wheel_balancer(balancer);
  if (!h || !state) return fail(UPKIE_B200_EINVAL, "wheel_balancer_state: invalid argument");
  CUDA_TRY(cudaSetDevice(h->device));
  // [4][n] -> [n][4]
  for (int k = 0; k < 4; ++k)
    CUDA_TRY(cudaMemcpy2DAsync(state + k, 4 * sizeof(float), h->state + size_t(k) * h->n, sizeof(float), sizeof(float), h->n,
                               cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return UPKIE_B200_OK;
}

int upkie_b200_observers_create(const UpkieObserverConfig* config, int n_robots, int device, void** observers) {
  if (!config || !observers) return fail(UPKIE_B200_EINVAL, "observers_create: null argument");
  std::string err;
  int rc = observers_create_impl(*config, n_robots, device, observers, err);
  return rc ? fail(rc, err) : UPKIE_B200_OK;
}
void upkie_b200_observers_destroy(void* observers) { observers_destroy_impl(observers); }
int upkie_b200_observers_reset(void* observers, const uint8_t* mask, void* stream) {
  std::string err;
  int rc = observers_reset_impl(observers, mask, static_cast<cudaStream_t>(stream), err);
  return rc ? fail(rc, err) : UPKIE_B200_OK;
}
int upkie_b200_observers_step(void* observers, const float* spine_obs, float* out, void* stream) {
  std::string err;
  int rc = observers_step_impl(observers, spine_obs, out, static_cast<cudaStream_t>(stream), err);
  return rc ? fail(rc, err) : UPKIE_B200_OK;
}

}  // extern "C"

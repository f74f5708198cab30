// SPDX-License-Identifier: Apache-2.0
//
// step_kernel.cuh -- the env-step kernel (one thread = one robot) and its launcher template.
//   k_step<MODE, AUTORESET, NOISE, TILE>   one 5 ms env tick: action front-end, 5 x (moteus torque law,
//       articulated-body dynamics, wheel-ground contact solve, semi-implicit integration), observation,
//       termination; optional fused auto-reset; optional torque noise models.
//       NOISE: 0 plain, 1 "extras" (noise models, external forces), 2 extras + joint-limit rows, 3 = 2 + spine timing
//       (+ body-ground contact rows), 4 = 2 + body-ground contact rows.
// Included by step_device.cu (TILE=0), step_host.cu (TILE=1), step_multicast.cu (TILE=2) and the two *_limits.cu
// units (NOISE=2), see kernel_common.cuh.
#pragma once

#include "kernel_common.cuh"

// 1 in step_device_limits.cu / step_host_limits.cu: that translation unit holds the NOISE=2 instantiations only
// (separate units so that the build stays parallel and the other units' kernels are untouched)
#ifndef UPKIE_STEP_LIMITS_TU
#define UPKIE_STEP_LIMITS_TU 0
#endif
#ifndef UPKIE_STEP_SPINE_TU
#define UPKIE_STEP_SPINE_TU 0  // 1 in step_*_spine.cu: the NOISE=3 instantiations (extras + limits + spine timing), UpkieServos
#endif
#ifndef UPKIE_STEP_BODY_TU
#define UPKIE_STEP_BODY_TU 0  // 1 in step_*_body.cu: the NOISE=4 instantiations (extras + limits + body-ground contact rows)
#endif
#ifndef UPKIE_ACTION_IN_TILE
#define UPKIE_ACTION_IN_TILE 0  // build-time experiment (tools/variants.py)
#endif

namespace upkie_b200 {
namespace {

// NVSwitch multicast stores (TILE=2): one store to the multicast address of a symmetric buffer is replicated by the
// switch into the same offset of every GPU's buffer -- the rollout "all-gather" without any collective kernel or copy.
__device__ __forceinline__ void mc_store4(float4* addr, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mc_store_u32(uint32_t* addr, uint32_t v) {
  asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}

// ---- one env tick of the robot `tid` --------------------------------------------------
// `tile4` is this warp's staging tile (TILE=1): on entry it holds the warp's 32 action rows when
// `full` (prefetched by the caller), and it is reused to transpose the observation rows on the way out.
template <int MODE, int AUTORESET, int NOISE, int TILE>
__device__ __forceinline__ void step_env(
    const SimParams& P, int tid, int n, int n_pad, float* __restrict__ state, const float* __restrict__ action,
    float* __restrict__ obs, float* __restrict__ reward, uint8_t* __restrict__ terminated,
    uint8_t* __restrict__ truncated, const float* __restrict__ eps_all, const float* __restrict__ mu_all,
    uint32_t* __restrict__ err, uint8_t* __restrict__ done_prev, uint32_t* __restrict__ episode,
    uint32_t* __restrict__ tick, uint64_t seed, uint64_t env_offset, const float* __restrict__ ext,
    uint32_t ext_local, float4* tile4, bool full, bool compact, const PeerPtrs* peers = nullptr,
    float* __restrict__ lag = nullptr) {
  const bool live = tid < n;
  const int i = live ? tid : n - 1;  // tail lanes shadow the last robot, stores masked
  const int lane = threadIdx.x & 31;
  const int wb = tid - lane;  // first env of this warp

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

  bool resetting = false;
  if (AUTORESET == AUTORESET_NEXT_STEP) resetting = done_prev[i] != 0;

  uint32_t e = 0;
  float a[UPKIE_ACT_DIM];
  float a0 = 0.f, a1 = 0.f;
  if (MODE == MODE_SERVOS) {
    if (TILE && full) {
#pragma unroll
      for (int k = 0; k < UPKIE_ACT_DIM / 4; ++k) {
        const float4 v = tile4[lane * (UPKIE_ACT_DIM / 4) + k];
        a[4 * k + 0] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
      }
      __syncwarp();
    } else {
      const float4* ap = reinterpret_cast<const float4*>(action + size_t(i) * UPKIE_ACT_DIM);
#pragma unroll
      for (int k = 0; k < UPKIE_ACT_DIM / 4; ++k) {
        const float4 v = __ldg(ap + k);
        a[4 * k + 0] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
      }
    }
  } else if (MODE == MODE_GYROPOD) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(action) + i);
    a0 = v.x; a1 = v.y;
  } else {
    a0 = __ldg(action + i);
    a1 = 0.f;  // upkie_pendulum.py:137
  }

  // One inlined copy of the physics serves both the regular tick (nb_substeps
  // substeps under the torque law) and the fused auto-reset (new initial state,
  // ONE zero-torque substep, pybullet_backend.py:227-228): resetting lanes run a
  // single iteration of the same loop, which keeps the kernel's code small
  // (instruction-cache footprint) and the warp converged.
  NoiseCtx nz{env_offset + uint64_t(i), 0u};
  if (NOISE) {
    nz.tick = tick[i] + 1u;
    if (live) tick[i] = nz.tick;
  }
  // external forces ride on the NOISE ("extras") instantiations so the plain kernels stay untouched
  const ExtForces xf{ext ? ext + i : nullptr, size_t(n_pad), ext_local};
  int nsub = P.nb_substeps;
  // spine mode (config.spine_mode, UpkieServos, the NOISE=2 kernels): every substep is one cycle of the Bullet spine
  // (NOISE = 3: its own instantiations, step_*_spine.cu, so that the other kernels do not carry the lag record)
  constexpr bool spine = NOISE == 3 && MODE == MODE_SERVOS;
  SpineLag L;
  if (spine) {
    float lr[UPKIE_LAG_DIM];
#pragma unroll
    for (int k = 0; k < UPKIE_LAG_DIM; ++k) lr[k] = lag[size_t(k) * n_pad + i];
    lag_from_row(lr, L);
  }
  if (resetting) {
    const uint32_t ep = episode[i] + 1u;
    if (live) episode[i] = ep;
    float init[UPKIE_INIT_DIM];
    sample_init_state(P, seed, env_offset + uint64_t(i), uint64_t(ep), init);
    if (spine) {
      reset_pose_spine(P, S, init, L);
      nsub = 3;  // three cycles with the servos stopped (Spine.cpp:119-125)
    } else {
      reset_pose(S, init);
      nsub = 1;
    }
  } else {
    if (MODE != MODE_SERVOS) e |= gyropod_action(P, S, a0, a1, a);
    e |= clamp_servo_action(P, a);
  }
#if UPKIE_ACTION_IN_TILE
  // TILE kernels: the clamped action row goes back to the warp's staging tile and is re-read at the top of every
  // substep (9 conflict-free LDS.128), instead of living in 36 registers / local-memory slots across the substep body
  const bool a_in_tile = TILE && full && MODE == MODE_SERVOS;
  if (a_in_tile) {
#pragma unroll
    for (int k = 0; k < UPKIE_ACT_DIM / 4; ++k)
      tile4[lane * (UPKIE_ACT_DIM / 4) + k] = make_float4(a[4 * k], a[4 * k + 1], a[4 * k + 2], a[4 * k + 3]);
    __syncwarp();
  }
#endif
  if (spine && !resetting) spine_assemble_observation(S, L);  // the first cycle's observation (Spine.cpp:126-131)
  const int nloop = (AUTORESET == AUTORESET_NEXT_STEP && spine && P.nb_substeps < 3) ? 3 : P.nb_substeps;
  for (int sub = 0; sub < nloop; ++sub) {
#if UPKIE_PHASE_SYNC_LEVEL >= 1
    __syncthreads();  // once per substep: all threads are converged here
#endif
#if UPKIE_ACTION_IN_TILE
    if (a_in_tile) {
#pragma unroll
      for (int k = 0; k < UPKIE_ACT_DIM / 4; ++k) {
        const float4 v = tile4[lane * (UPKIE_ACT_DIM / 4) + k];
        a[4 * k + 0] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
      }
    }
#endif
    if (sub < nsub) {
      // the body-ground contacts of the tick's last substep go to the handle's record (NOISE >= 3 kernels)
      const BodyRecOut br{(NOISE >= 3 && P.body_rec && live && sub == nsub - 1) ? P.body_rec + i : nullptr,
                          size_t(P.body_rec_stride)};
      if (spine) {
        if (resetting && sub == 2) spine_assemble_observation(S, L);
        spine_cycle(P, S, L, a, resetting, eps, mu, WarpAny(), PhaseSync(), P.joint_limits >= 1 ? P.joint_limits : 1, br);
      } else {
        servo_substep(P, S, a, resetting, eps, mu, WarpAny(), PhaseSync(), NOISE ? &nz : nullptr, sub,
                      (NOISE && ext) ? &xf : nullptr, NOISE >= 2 ? (P.joint_limits >= 1 ? P.joint_limits : 1) : 0, br);
      }
    } else {
#pragma unroll
      for (int k = 0; k < kPhaseSyncs; ++k) PhaseSync()();
    }
  }
  if (!spine) observe_update(P, S);  // spine mode: the cycles read the IMU
  if (resetting) {
    reset_wrapper_state(S);
  } else {
    e |= state_sanity(S);
    if (MODE != MODE_SERVOS) {
      S.yaw += a1 * P.dt;  // integrates the unclamped action[1], upkie_gyropod.py:383-385
      S.yaw_vel = a1;
    }
  }

  // observation, reward, termination
  bool term = false;
  float o6[6];
  if (MODE == MODE_SERVOS) {
    if (P.servos_fall_termination) term = (fabsf(base_pitch(S)) > P.fall_pitch) || (S.pos[2] < P.min_base_height);
  } else {
    gyropod_obs(P, S, o6);
    term = fabsf(o6[1]) > P.fall_pitch;  // strict, upkie_gyropod.py:345
  }
  if (resetting) term = false;

  if (AUTORESET == AUTORESET_SAME_STEP) {
    if (term) {
      const uint32_t ep = episode[i] + 1u;
      if (live) episode[i] = ep;
      float init[UPKIE_INIT_DIM];
      sample_init_state(P, seed, env_offset + uint64_t(i), uint64_t(ep), init);
      const BodyRecOut br{(NOISE >= 3 && P.body_rec && live) ? P.body_rec + i : nullptr, size_t(P.body_rec_stride)};
      if (spine) reset_robot_spine(P, S, L, init, eps, mu, WarpAny(), P.joint_limits, br);
      else reset_robot(P, S, init, eps, mu, WarpAny(), NOISE >= 2 ? P.joint_limits : 0, br);
      if (MODE != MODE_SERVOS) gyropod_obs(P, S, o6);
    }
  }

  if (live) store_state(state, n_pad, i, S);
  if (spine && live) {
    float lr[UPKIE_LAG_DIM];
    lag_to_row(L, lr);
#pragma unroll
    for (int k = 0; k < UPKIE_LAG_DIM; ++k) lag[size_t(k) * n_pad + i] = lr[k];
  }
  if (MODE == MODE_SERVOS) {
    float o[UPKIE_OBS_DIM];
    float tq[6];
    measured_torques(P, S, NOISE ? &nz : nullptr, tq);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      o[j * 5 + 0] = S.q[j]; o[j * 5 + 1] = S.qd[j]; o[j * 5 + 2] = tq[j];
      o[j * 5 + 3] = 42.0f;  // pybullet_backend.py:471
      o[j * 5 + 4] = 18.0f;  // pybullet_backend.py:472
    }
    if (spine) {  // the replies of two cycles ago, as the spine's observation reports them
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        o[j * 5 + 0] = L.obs_rep[3 * j]; o[j * 5 + 1] = L.obs_rep[3 * j + 1]; o[j * 5 + 2] = L.obs_rep[3 * j + 2];
        o[j * 5 + 3] = 20.0f;  // BulletInterface.cpp:70
      }
    }
    if (TILE && compact) {
      // compact rows [6 joints][position, velocity, torque]: temperature / voltage are constants the
      // host fills once (pybullet_backend.py:471-472), 72 B per env instead of 120 B over PCIe
      float c[18];
#pragma unroll
      for (int j = 0; j < 6; ++j) { c[3 * j] = o[5 * j]; c[3 * j + 1] = o[5 * j + 1]; c[3 * j + 2] = o[5 * j + 2]; }
      if (full) {
        float2* t2 = reinterpret_cast<float2*>(tile4) + lane * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) t2[k] = make_float2(c[2 * k], c[2 * k + 1]);
        __syncwarp();
        float4* op = reinterpret_cast<float4*>(obs + size_t(wb) * 18);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int idx = k * 32 + lane;
          if (idx < 32 * 18 / 4) {
            if (TILE == 2 && !peers->deferred) {
              const float4 v = tile4[idx];
              if (peers->n == 0) {
                mc_store4(op + idx, v);
              } else {
                // `obs` is unused here: the row offset of this step's slot is already in every peer pointer
                for (int p = 0; p < peers->n; ++p) reinterpret_cast<float4*>(peers->obs[p] + size_t(wb) * 18)[idx] = v;
              }
            } else {
              op[idx] = tile4[idx];
            }
          }
        }
      } else if (live) {
        float2* op = reinterpret_cast<float2*>(obs + size_t(i) * 18);
#pragma unroll
        for (int k = 0; k < 9; ++k) op[k] = make_float2(c[2 * k], c[2 * k + 1]);
      }
    } else if (TILE && full) {
      float2* t2 = reinterpret_cast<float2*>(tile4) + lane * (UPKIE_OBS_DIM / 2);
#pragma unroll
      for (int k = 0; k < UPKIE_OBS_DIM / 2; ++k) t2[k] = make_float2(o[2 * k], o[2 * k + 1]);
      __syncwarp();
      float4* op = reinterpret_cast<float4*>(obs + size_t(wb) * UPKIE_OBS_DIM);
#pragma unroll
      for (int k = 0; k < (32 * UPKIE_OBS_DIM / 4 + 31) / 32; ++k) {
        const int idx = k * 32 + lane;
        if (idx < 32 * UPKIE_OBS_DIM / 4) op[idx] = tile4[idx];
      }
    } else if (live) {
      float2* op = reinterpret_cast<float2*>(obs + size_t(i) * UPKIE_OBS_DIM);
#pragma unroll
      for (int k = 0; k < UPKIE_OBS_DIM / 2; ++k) op[k] = make_float2(o[2 * k], o[2 * k + 1]);
    }
  } else if (MODE == MODE_GYROPOD) {
    if (TILE && full) {
      float2* t2 = reinterpret_cast<float2*>(tile4) + lane * 3;
      t2[0] = make_float2(o6[0], o6[1]);
      t2[1] = make_float2(o6[2], o6[3]);
      t2[2] = make_float2(o6[4], o6[5]);
      __syncwarp();
      float4* op = reinterpret_cast<float4*>(obs + size_t(wb) * 6);
      op[lane] = tile4[lane];
      if (lane < 16) op[32 + lane] = tile4[32 + lane];
    } else if (live) {
      float2* op = reinterpret_cast<float2*>(obs + size_t(i) * 6);
      op[0] = make_float2(o6[0], o6[1]);
      op[1] = make_float2(o6[2], o6[3]);
      op[2] = make_float2(o6[4], o6[5]);
    }
  } else if (live) {
    // upkie_pendulum.py:17 _PENDULUM_OBS_INDICES = [1, 0, 4, 3]
    reinterpret_cast<float4*>(obs)[i] = make_float4(o6[1], o6[0], o6[4], o6[3]);
  }
  if (TILE == 2) {
    // the host only launches this variant on full, aligned warps (n % 32 == 0): the warp's 32 `terminated` bytes go
    // out as eight words built from the ballot (multimem.st has no byte form)
    const unsigned m = __ballot_sync(0xffffffffu, term);
    if (lane < 8) {
      const unsigned nib = (m >> (4 * lane)) & 0xFu;
      const uint32_t word = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
      if (peers->deferred) {
        reinterpret_cast<uint32_t*>(terminated + wb)[lane] = word;  // local slot; sent by a later launch's prologue
      } else if (peers->n == 0) {
        mc_store_u32(reinterpret_cast<uint32_t*>(terminated + wb) + lane, word);
      } else {
        for (int p = 0; p < peers->n; ++p) reinterpret_cast<uint32_t*>(peers->term[p] + wb)[lane] = word;
      }
    }
  }
  if (!live) return;
  if (reward) reward[i] = 0.0f;  // upkie_env.py:230
  if (TILE != 2) terminated[i] = term ? 1 : 0;
  if (truncated) truncated[i] = 0;
  if (e) err[i] |= e;
  if (AUTORESET == AUTORESET_NEXT_STEP) done_prev[i] = term ? 1 : 0;
}

// Deferred rollout transport: send the warp's 32 compact rows (and `terminated` bytes) of an EARLIER step, read from
// this rank's local slot of that step, to every GPU - NVSwitch multicast store or plain stores into the peers'
// buffers. Called at the top of a tile, before its physics: the stores drain while the warp simulates.
__device__ __forceinline__ void push_rows(const PeerPtrs& pp, int wb, int lane) {
  const float4* src = reinterpret_cast<const float4*>(pp.src_obs + size_t(wb) * 18);
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int idx = k * 32 + lane;
    if (idx < 32 * 18 / 4) {
      const float4 v = src[idx];
      if (pp.n == 0) {
        mc_store4(reinterpret_cast<float4*>(pp.mc_obs + size_t(wb) * 18) + idx, v);
      } else {
        for (int p = 0; p < pp.n; ++p) reinterpret_cast<float4*>(pp.obs[p] + size_t(wb) * 18)[idx] = v;
      }
    }
  }
  if (lane < 8) {
    const uint32_t word = reinterpret_cast<const uint32_t*>(pp.src_term + wb)[lane];
    if (pp.n == 0) {
      mc_store_u32(reinterpret_cast<uint32_t*>(pp.mc_term + wb) + lane, word);
    } else {
      for (int p = 0; p < pp.n; ++p) reinterpret_cast<uint32_t*>(pp.term[p] + wb)[lane] = word;
    }
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- the env-step kernel ------------------------------------------------------------
// TILE=0 (device buffers): one env per thread, grid = ceil(cnt / block).
// TILE=1 (host buffers, zero-copy): persistent blocks walk the tiles of `blockDim.x` envs with stride
// gridDim.x. The action rows of the NEXT tile are fetched over PCIe with cp.async into the second
// shared-memory buffer while the current tile computes, and observation rows leave through coalesced
// 16 B stores, so the link streams in both directions for the whole launch instead of in bursts
// between compute phases.
template <int MODE, int AUTORESET, int NOISE, int TILE>
#ifdef UPKIE_MAXNREG
__global__ void __maxnreg__(UPKIE_MAXNREG)
#else
__global__ void __launch_bounds__(UPKIE_MAX_THREADS, UPKIE_MIN_BLOCKS)
#endif
k_step(const __grid_constant__ SimParams P, int i0, int n, int n_pad, float* __restrict__ state,
       const float* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
       uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated, const float* __restrict__ eps_all,
       const float* __restrict__ mu_all, uint32_t* __restrict__ err, uint8_t* __restrict__ done_prev,
       uint32_t* __restrict__ episode, uint32_t* __restrict__ tick, uint64_t seed, uint64_t env_offset,
       const float* __restrict__ ext, uint32_t ext_local, int coalesce, const __grid_constant__ PeerPtrs peers,
       float* __restrict__ lag) {
  // this launch covers the envs [i0, n)
  if (!TILE) {
    step_env<MODE, AUTORESET, NOISE, 0>(P, i0 + blockIdx.x * blockDim.x + threadIdx.x, n, n_pad, state, action, obs,
                                        reward, terminated, truncated, eps_all, mu_all, err, done_prev, episode, tick,
                                        seed, env_offset, ext, ext_local, nullptr, false, false, nullptr, lag);
    return;
  }
  extern __shared__ float4 s_tile[];
  constexpr int kRow4 = 32 * UPKIE_ACT_DIM / 4;  // float4 per warp tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float4* buf[2] = {s_tile + warp * kRow4, s_tile + (nwarps + warp) * kRow4};
  const int ntiles = (n - i0 + blockDim.x - 1) / blockDim.x;
  // a warp's rows are prefetched when all 32 envs exist and the rows are 16 B aligned
  auto warp_full = [&](int t) { return (coalesce & 1) && (i0 + t * int(blockDim.x) + warp * 32 + 32 <= n); };
  auto prefetch = [&](int t, float4* dst) {
    if (MODE == MODE_SERVOS && warp_full(t)) {
      const float4* src =
          reinterpret_cast<const float4*>(action + size_t(i0 + t * int(blockDim.x) + warp * 32) * UPKIE_ACT_DIM);
#pragma unroll
      for (int k = 0; k < UPKIE_ACT_DIM / 4; ++k) cp_async16(dst + k * 32 + lane, src + k * 32 + lane);
    }
    cp_async_commit();
  };
  int t = blockIdx.x, it = 0;
  if (t < ntiles) prefetch(t, buf[0]);
  for (; t < ntiles; t += gridDim.x, ++it) {
    const int nt = t + gridDim.x;
    if (nt < ntiles) {
      prefetch(nt, buf[(it + 1) & 1]);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    if (TILE == 2) {
      if (peers.deferred && peers.src_obs) push_rows(peers, i0 + t * int(blockDim.x) + warp * 32, lane);
    }
    step_env<MODE, AUTORESET, NOISE, TILE>(P, i0 + t * blockDim.x + threadIdx.x, n, n_pad, state, action, obs, reward,
                                        terminated, truncated, eps_all, mu_all, err, done_prev, episode, tick, seed,
                                        env_offset, ext, ext_local, buf[it & 1], warp_full(t), (coalesce & 2) != 0,
                                        TILE == 2 ? &peers : nullptr, lag);
    __syncwarp();  // the tile is free again before the next prefetch lands in it
  }
}


template <int TILE, int MODE>
cudaError_t launch_step_mode(const StepArgs& a) {
  const int tiles = (a.cnt + a.block - 1) / a.block;
  const int grid = (TILE && a.grid > 0 && a.grid < tiles) ? a.grid : tiles;
  const size_t smem = TILE ? size_t(2) * (a.block / 32) * 32 * UPKIE_ACT_DIM * sizeof(float) : 0;
  // the tile path needs 16 B aligned rows of 32 envs
  const int aligned =
      ((reinterpret_cast<uintptr_t>(a.action) | reinterpret_cast<uintptr_t>(a.obs)) & 15) == 0 && (a.i0 % 32) == 0;
  const int coalesce = (aligned ? 1 : 0) | ((TILE && a.compact_obs) ? 2 : 0);  // bit 0 tile path, bit 1 compact rows
#define LAUNCH_N(AR, NZ)                                                                                         \
  k_step<MODE, AR, NZ, TILE><<<grid, a.block, smem, a.stream>>>(                                                 \
      *a.P, a.i0, a.i0 + a.cnt, a.n_pad, a.state, a.action, a.obs, a.reward, a.terminated, a.truncated, a.eps,   \
      a.mu, a.err, a.done_prev, a.episode, a.tick, a.seed, a.env_offset, a.ext, a.ext_local, coalesce, a.peers, a.lag)
#if UPKIE_STEP_SPINE_TU
#define LAUNCH(AR) LAUNCH_N(AR, 3)
#elif UPKIE_STEP_BODY_TU
#define LAUNCH(AR) LAUNCH_N(AR, 4)
#elif UPKIE_STEP_LIMITS_TU
#define LAUNCH(AR) LAUNCH_N(AR, 2)
#else
#define LAUNCH(AR)                \
  do {                            \
    if (a.noise) LAUNCH_N(AR, 1); \
    else LAUNCH_N(AR, 0);         \
  } while (0)
#endif
  if (a.autoreset == AUTORESET_NEXT_STEP) LAUNCH(AUTORESET_NEXT_STEP);
  else if (a.autoreset == AUTORESET_SAME_STEP) LAUNCH(AUTORESET_SAME_STEP);
  else LAUNCH(AUTORESET_DISABLED);
#undef LAUNCH
#undef LAUNCH_N
  return cudaGetLastError();
}

template <int TILE>
cudaError_t launch_step_kernels(const StepArgs& a) {
  if (a.mode == MODE_SERVOS) return launch_step_mode<TILE, MODE_SERVOS>(a);
  if (a.mode == MODE_GYROPOD) return launch_step_mode<TILE, MODE_GYROPOD>(a);
  return launch_step_mode<TILE, MODE_PENDULUM>(a);
}

}  // namespace
}  // namespace upkie_b200

// SPDX-License-Identifier: Apache-2.0
//
// kernel_common.cuh -- what the translation units of libupkie_b200.so share: build-time knobs, the
// state load/store helpers, and the launch descriptor of the env-step kernel. The step kernel itself
// (step_kernel.cuh) is instantiated in two translation units that nvcc compiles in parallel:
//   step_device.cu  TILE=0  per-thread action loads / observation stores (device buffers)
//   step_host.cu    TILE=1  per-warp shared-memory tile with coalesced 16 B accesses, the variant whose
//                           warps read actions from and write observations to mapped pinned HOST memory
// Keeping them apart leaves the register allocation of the device-buffer kernel untouched (the tile
// costs it 8 % when both paths live in one kernel, profiles/r01_variants.md).
#pragma once

#include <cuda_runtime.h>

#include "params.h"

namespace upkie_b200 {

// build-time tuning knobs (tools/variants.py explores them; defaults are the measured best)
#ifndef UPKIE_MAX_THREADS
#define UPKIE_MAX_THREADS 256
#endif
#ifndef UPKIE_MIN_BLOCKS
#define UPKIE_MIN_BLOCKS 1
#endif
#ifndef UPKIE_DEFAULT_BLOCK
#define UPKIE_DEFAULT_BLOCK 0  // 0 = pick per launch (pick_block)
#endif
#ifndef UPKIE_PHASE_SYNC_LEVEL  // 0 none, 1 one barrier per substep, 2 also six barriers inside the substep
#define UPKIE_PHASE_SYNC_LEVEL 1
#endif

enum { MODE_SERVOS = 0, MODE_GYROPOD = 1, MODE_PENDULUM = 2 };
enum { AUTORESET_DISABLED = 0, AUTORESET_NEXT_STEP = 1, AUTORESET_SAME_STEP = 2 };

// CTA-wide barrier that tolerates intra-warp divergence (non-.aligned form): every
// thread of the block arrives exactly kPhaseSyncs times per substep. It buys no
// data exchange: it keeps the block's warps within the same instruction-cache
// window of the ~100 KB substep body (ncu: stall_no_instruction was the top stall).
struct PhaseSync {
  __device__ __forceinline__ void operator()() const {
#if UPKIE_PHASE_SYNC_LEVEL >= 2
    asm volatile("barrier.sync 0;" ::: "memory");
#endif
  }
};

struct WarpAny {
  __device__ __forceinline__ bool operator()(bool p) const { return __any_sync(__activemask(), p); }
};

__device__ __forceinline__ void load_state(const float* __restrict__ st, int n_pad, int i, RobotState& S) {
  float r[UPKIE_STATE_DIM];
#pragma unroll
  for (int k = 0; k < UPKIE_STATE_DIM; ++k) r[k] = st[size_t(k) * n_pad + i];
  state_from_row(r, S);
}

__device__ __forceinline__ void store_state(float* __restrict__ st, int n_pad, int i, const RobotState& S) {
  float r[UPKIE_STATE_DIM];
  state_to_row(S, r);
#pragma unroll
  for (int k = 0; k < UPKIE_STATE_DIM; ++k) st[size_t(k) * n_pad + i] = r[k];
}

// TILE=2 (rollout rows leave the GPU from inside the step kernel): n = 0 -> `obs` / `terminated` are NVSwitch multicast
// addresses (multimem.st, the switch replicates the store into every GPU's buffer); n > 0 -> plain stores into the n
// peer buffers listed here (peer-mapped symmetric memory over NVLink; the list includes this rank's own buffer).
//
// `deferred` (round 2, the default of bench.py): the rows this launch sends are those of an EARLIER step - read from
// `src_obs` / `src_term` (this rank's local slot of that step) and sent in the PROLOGUE of the kernel, so that their
// NVLink latency hides under the ~0.1 ms of simulation instead of holding up the completion of the launch - while this
// step's rows go to the local slot (`obs` / `terminated` of the launch) with plain stores.
struct PeerPtrs {
  float* obs[UPKIE_MAX_PEERS];
  uint8_t* term[UPKIE_MAX_PEERS];
  int n;
  int deferred;
  float* mc_obs;          // deferred + multicast: multicast address of the earlier step's slot (null with peer stores)
  uint8_t* mc_term;
  const float* src_obs;   // deferred: local rows of the earlier step, null = nothing to send in this launch
  const uint8_t* src_term;
};

// One launch of the env-step kernel over the envs [i0, i0 + cnt) of a handle.
struct StepArgs {
  const SimParams* P;
  int mode, autoreset, noise;  // noise: 1 = the "extras" instantiation (torque noise models, external forces), 2 = extras + joint-limit rows, 3 = 2 + spine timing, 4 = 2 + body-ground contact rows
  int i0, cnt, n_pad, block;
  int compact_obs;      // TILE=1, servos: observation rows [6][3] (position, velocity, torque) instead of [6][5]
  int grid;             // TILE=1: number of persistent blocks (0 = one block per tile)
  float* state;
  const float* action;
  float* obs;
  float* reward;        // may be null
  uint8_t* terminated;
  uint8_t* truncated;   // may be null
  const float* eps;
  const float* mu;
  uint32_t* err;
  uint8_t* done_prev;
  uint32_t* episode;
  uint32_t* tick;
  uint64_t seed, env_offset;
  const float* ext;     // external forces [21][n_pad] or null (noise != 0 when set)
  uint32_t ext_local;
  cudaStream_t stream;
  PeerPtrs peers;       // TILE=2 only
  float* lag;           // spine mode: [UPKIE_LAG_DIM][n_pad] lag records, else null
};

cudaError_t launch_step_device(const StepArgs& a);  // step_device.cu
cudaError_t launch_step_host(const StepArgs& a);    // step_host.cu
cudaError_t launch_step_multicast(const StepArgs& a);  // step_multicast.cu: TILE=2, UpkieServos, compact rows
cudaError_t launch_step_device_limits(const StepArgs& a);  // step_device_limits.cu: NOISE=2 (joint-limit rows), TILE=0
cudaError_t launch_step_host_limits(const StepArgs& a);    // step_host_limits.cu: NOISE=2, TILE=1
cudaError_t launch_step_multicast_limits(const StepArgs& a);  // step_multicast_limits.cu: NOISE=2, TILE=2
cudaError_t launch_push_rows(const PeerPtrs& pp, int n, cudaStream_t stream);  // step_multicast.cu: flush of a rollout's last rows
cudaError_t launch_step_device_body(const StepArgs& a);  // step_device_body.cu: NOISE=4 (limits + body-ground contact rows), TILE=0
cudaError_t launch_step_host_body(const StepArgs& a);    // step_host_body.cu: NOISE=4, TILE=1
cudaError_t launch_step_device_spine(const StepArgs& a);  // step_device_spine.cu: NOISE=3 (spine timing), TILE=0
cudaError_t launch_step_host_spine(const StepArgs& a);    // step_host_spine.cu: NOISE=3, TILE=1

}  // namespace upkie_b200

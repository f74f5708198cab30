# SPDX-License-Identifier: Apache-2.0
"""Tensor-level front end of the vectorised simulation (one handle = one GPU).

``UpkieSim`` owns a ``libupkie_b200`` handle and exposes the flat fast path of
SURVEY.md section 8(b): ``step_servos(action[N, 6, 6]) -> obs[N, 6, 5], reward[N],
terminated[N], truncated[N]`` over PyTorch CUDA tensors (PyTorch is used for
device memory and streams only; all arithmetic happens in the sm_100a kernels).
"""

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _abi
from ._lib import check, lib
from .exceptions import UpkieRuntimeError
from .model import Model, default_model

AUTORESET_DISABLED, AUTORESET_NEXT_STEP, AUTORESET_SAME_STEP = 0, 1, 2


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class UpkieSim:
    """N independent robots on one CUDA device.

    Replaces ``PyBulletBackend`` (``upkie/envs/backends/pybullet_backend.py:31``)
    for N robots at once; same constructor knobs (``dt``, ``nb_substeps``,
    ``torque_control_kp/kd``, ``joint_properties`` friction, ``inertia_variation``)
    through ``config`` / ``set_randomization``.
    """

    def __init__(
        self,
        n_envs: int,
        model: Optional[Model] = None,
        config: Optional[_abi.UpkieSimConfig] = None,
        device: int = 0,
    ):
        if not torch.cuda.is_available():
            raise UpkieRuntimeError(
                "upkie_b200 needs a CUDA device (there is no CPU fallback)"
            )
        self.model = model if model is not None else default_model()
        self.config = config if config is not None else _abi.default_sim_config()
        self.n = int(n_envs)
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self._model_struct = self.model.to_struct()
        self._h = C.c_void_p()
        check(
            lib().upkie_b200_create(
                C.byref(self._model_struct), C.byref(self.config), self.n, self.device_index, C.byref(self._h)
            )
        )
        f32, u8 = torch.float32, torch.uint8
        dev = self.device
        self.reward = torch.empty(self.n, dtype=f32, device=dev)
        self.terminated = torch.empty(self.n, dtype=u8, device=dev)
        self.truncated = torch.empty(self.n, dtype=u8, device=dev)
        self.obs_servos = torch.empty((self.n, 6, 5), dtype=f32, device=dev)
        self.obs_gyropod = torch.empty((self.n, 6), dtype=f32, device=dev)
        self.obs_pendulum = torch.empty((self.n, 4), dtype=f32, device=dev)

    # ------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().upkie_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_tensor(self, t: torch.Tensor, shape, dtype=torch.float32, name="tensor"):
        if t.device != self.device or t.dtype != dtype or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
            raise UpkieRuntimeError(
                f"{name}: expected contiguous {dtype} tensor of shape {tuple(shape)} on {self.device}, "
                f"got {t.dtype} {tuple(t.shape)} on {t.device}"
            )

    # ------------------------------------------------------------------
    def set_autoreset(self, mode: int, seed: int = 0, env_offset: int = 0) -> None:
        check(lib().upkie_b200_set_autoreset(self._h, int(mode), int(seed), int(env_offset)))
        self._autoreset = (int(mode), int(seed), int(env_offset))

    def set_config(self, config: _abi.UpkieSimConfig) -> None:
        """Replace the configuration of the live handle (``upkie_b200_set_config``): initial-state bounds of the
        on-device reset sampler, noise levels, gains. Steps enqueued afterwards use it."""
        check(lib().upkie_b200_set_config(self._h, C.byref(config)))
        self.config = config

    def set_randomization(self, friction: Optional[torch.Tensor] = None, inertia_eps: Optional[torch.Tensor] = None):
        """Per-env floor friction [N] and ``randomize_inertias`` epsilons [N, 6]
        (``pybullet_backend.py:571-601``)."""
        if friction is not None:
            self._check_tensor(friction, (self.n,), name="friction")
        if inertia_eps is not None:
            self._check_tensor(inertia_eps, (self.n, 6), name="inertia_eps")
        check(lib().upkie_b200_set_randomization(self._h, _ptr(friction), _ptr(inertia_eps), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()
        self._randomization = (None if friction is None else friction.clone(), None if inertia_eps is None else inertia_eps.clone())

    def set_external_forces(self, force: Optional[torch.Tensor] = None, local_mask: int = 0) -> None:
        """``force[N, 7, 3]`` newtons at the centres of mass of the 7 bodies, applied on every substep of
        the following steps until overwritten; ``None`` clears. Bit ``b`` of ``local_mask``: the force on
        body ``b`` is expressed in the body frame (``pybullet_backend.py:603-658``)."""
        if force is not None:
            self._check_tensor(force, (self.n, _abi.NB, 3), name="force")
        check(lib().upkie_b200_set_external_forces(self._h, _ptr(force), int(local_mask), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()
        self._external = (None if force is None else force.clone(), int(local_mask))

    def reset(
        self,
        mask: Optional[torch.Tensor] = None,
        init_state: Optional[torch.Tensor] = None,
        seed: int = 0,
        env_offset: int = 0,
    ) -> None:
        """``PyBulletBackend.reset`` for the envs selected by ``mask`` (all if
        None). ``init_state[N, 25]`` = sampled ``RobotState`` rows; None = sample
        on the device."""
        if mask is not None:
            self._check_tensor(mask, (self.n,), torch.uint8, "mask")
        if init_state is not None:
            self._check_tensor(init_state, (self.n, _abi.INIT_DIM), name="init_state")
        check(lib().upkie_b200_reset(self._h, _ptr(mask), _ptr(init_state), int(seed), int(env_offset), self._stream()))

    def _outputs(self, obs, default_obs, reward, terminated, truncated):
        """Caller-provided output tensors (e.g. ``RolloutBuffer.slot(t)``: the kernel then
        writes the rollout in place) or the handle's own."""
        obs = default_obs if obs is None else obs
        reward = self.reward if reward is None else reward
        terminated = self.terminated if terminated is None else terminated
        truncated = self.truncated if truncated is None else truncated
        if obs is not default_obs:
            if obs.numel() != default_obs.numel() or obs.data_ptr() % 16 != 0:
                raise UpkieRuntimeError("obs: wrong number of elements or not 16-byte aligned")
            self._check_tensor(obs, obs.shape, name="obs")
        if reward is not self.reward:
            self._check_tensor(reward, (self.n,), name="reward")
        if terminated is not self.terminated:
            self._check_tensor(terminated, (self.n,), torch.uint8, "terminated")
        if truncated is not self.truncated:
            self._check_tensor(truncated, (self.n,), torch.uint8, "truncated")
        return obs, reward, terminated, truncated

    def step_servos(self, action: torch.Tensor, obs: Optional[torch.Tensor] = None, reward=None, terminated=None,
                    truncated=None):
        self._check_tensor(action, (self.n, 6, 6), name="action")
        obs, reward, terminated, truncated = self._outputs(obs, self.obs_servos, reward, terminated, truncated)
        check(
            lib().upkie_b200_step_servos(
                self._h, _ptr(action), _ptr(obs), _ptr(reward), _ptr(terminated), _ptr(truncated), self._stream(),
            )
        )
        return obs, reward, terminated, truncated

    def step_servos_compact(self, action: torch.Tensor, obs: Optional[torch.Tensor] = None, terminated=None):
        """Device buffers, compact rows: ``obs[N, 6, 3]`` (position, velocity, torque) and ``terminated``; the
        constants of the reference (temperature, voltage, reward, truncated) are not written. What a rollout
        buffer gathered across GPUs carries (``RolloutBuffer(..., compact=True)``)."""
        self._check_tensor(action, (self.n, 6, 6), name="action")
        if obs is None:
            if getattr(self, "obs_servos_compact", None) is None:
                self.obs_servos_compact = torch.empty((self.n, 6, 3), dtype=torch.float32, device=self.device)
            obs = self.obs_servos_compact
        elif obs.numel() != self.n * 18 or obs.dtype != torch.float32 or not obs.is_contiguous() or obs.data_ptr() % 16:
            raise UpkieRuntimeError("obs: expected contiguous 16-byte aligned float32 [N, 6, 3]")
        terminated = self.terminated if terminated is None else terminated
        if terminated is not self.terminated:
            self._check_tensor(terminated, (self.n,), torch.uint8, "terminated")
        check(lib().upkie_b200_step_servos_compact(self._h, _ptr(action), _ptr(obs), _ptr(terminated), self._stream()))
        return obs, terminated

    def step_servos_multicast(self, action: torch.Tensor, obs_mc_ptr: int, terminated_mc_ptr: int) -> None:
        """Compact rows and ``terminated`` go to NVSwitch multicast addresses (``PeerRolloutBuffer.multicast_slot``): every GPU of the node receives them."""
        self._check_tensor(action, (self.n, 6, 6), name="action")
        check(lib().upkie_b200_step_servos_multicast(
            self._h, _ptr(action), C.c_void_p(int(obs_mc_ptr)), C.c_void_p(int(terminated_mc_ptr)), self._stream()))

    def step_servos_peers(self, action: torch.Tensor, obs_ptrs, terminated_ptrs) -> None:
        """Compact rows and ``terminated`` stored by the step kernel into every buffer of ``obs_ptrs`` /
        ``terminated_ptrs`` (device addresses of this step's slot in each peer's symmetric rollout buffer,
        ``PeerRolloutBuffer.peer_slots``): the rollout gather over plain NVLink stores, no collective kernel."""
        self._check_tensor(action, (self.n, 6, 6), name="action")
        k = len(obs_ptrs)
        oa = (C.c_void_p * k)(*[C.c_void_p(int(x)) for x in obs_ptrs])
        ta = (C.c_void_p * k)(*[C.c_void_p(int(x)) for x in terminated_ptrs])
        check(lib().upkie_b200_step_servos_peers(self._h, _ptr(action), oa, ta, k, self._stream()))

    def step_servos_push(self, action: torch.Tensor, obs_ptr: int, terminated_ptr: int, push=None) -> None:
        """Deferred rollout transport (``upkie_b200_step_servos_push``): this step's compact rows go to the local slot
        at ``obs_ptr`` / ``terminated_ptr``; ``push`` (an ``_abi.UpkiePush`` from ``PeerRolloutBuffer.push_descriptor``)
        names an earlier slot the same launch sends to every GPU in its prologue, or None."""
        self._check_tensor(action, (self.n, 6, 6), name="action")
        check(lib().upkie_b200_step_servos_push(
            self._h, _ptr(action), C.c_void_p(int(obs_ptr)), C.c_void_p(int(terminated_ptr)),
            C.byref(push) if push is not None else None, self._stream()))

    def push_rows(self, push) -> None:
        """Send one slot on its own (last step of a rollout): ``upkie_b200_push_rows``."""
        check(lib().upkie_b200_push_rows(self._h, C.byref(push), self._stream()))

    def step_gyropod(self, action: torch.Tensor, obs: Optional[torch.Tensor] = None, reward=None, terminated=None,
                     truncated=None):
        self._check_tensor(action, (self.n, 2), name="action")
        obs, reward, terminated, truncated = self._outputs(obs, self.obs_gyropod, reward, terminated, truncated)
        check(
            lib().upkie_b200_step_gyropod(
                self._h, _ptr(action), 2, _ptr(obs), _ptr(reward), _ptr(terminated), _ptr(truncated), self._stream(),
            )
        )
        return obs, reward, terminated, truncated

    def step_pendulum(self, action: torch.Tensor, obs: Optional[torch.Tensor] = None, reward=None, terminated=None,
                      truncated=None):
        self._check_tensor(action, (self.n, 1), name="action")
        obs, reward, terminated, truncated = self._outputs(obs, self.obs_pendulum, reward, terminated, truncated)
        check(
            lib().upkie_b200_step_gyropod(
                self._h, _ptr(action), 1, _ptr(obs), _ptr(reward), _ptr(terminated), _ptr(truncated), self._stream(),
            )
        )
        return obs, reward, terminated, truncated

    # host-buffer path (H2D + kernel + D2H inside the call) ---------------------
    def _host_buffers(self):
        """Pinned host arrays owned by the handle: ``host_action_buffer()`` (fill it
        in place for a staging-free H2D copy) and the outputs returned by the
        ``*_host`` calls (overwritten by the next call)."""
        if getattr(self, "_hb", None) is None:
            def pinned(shape, dtype):
                return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()

            self._hb = {
                "act36": pinned((self.n, 6, 6), torch.float32),
                "act2": pinned((self.n, 2), torch.float32),
                "act1": pinned((self.n, 1), torch.float32),
                "obs30": pinned((self.n, 6, 5), torch.float32),
                "obs18": pinned((self.n, 6, 3), torch.float32),
                "obs6": pinned((self.n, 6), torch.float32),
                "obs4": pinned((self.n, 4), torch.float32),
                "rew": pinned((self.n,), torch.float32),
                "term": pinned((self.n,), torch.uint8),
                "trunc": pinned((self.n,), torch.uint8),
            }
            self._hb["rew"][:] = 0.0  # upkie_env.py:230
            self._hb["trunc"][:] = 0  # upkie_env.py:197
        return self._hb

    def host_action_buffer(self, act_dim: int = 36) -> np.ndarray:
        """Pinned ``[N, 6, 6]`` / ``[N, 2]`` / ``[N, 1]`` action array to fill in place."""
        return self._host_buffers()[f"act{act_dim}"]

    def step_servos_host(self, action: np.ndarray):
        """``action[N, 6, 6]`` float32 on the host (ideally ``host_action_buffer()``).
        Returns pinned arrays that the next ``*_host`` call overwrites."""
        hb = self._host_buffers()
        a = action
        if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(action, dtype=np.float32)
        if a.size != self.n * 36:
            raise UpkieRuntimeError(f"action: expected {self.n * 36} float32 values, got {a.size}")
        obs, rew, term, trunc = hb["obs30"], hb["rew"], hb["term"], hb["trunc"]
        # reward and truncated are constants of the reference (upkie_env.py:197,230): not transported
        check(lib().upkie_b200_step_servos_host(self._h, a.ctypes.data, obs.ctypes.data, None, term.ctypes.data, None))
        return obs, rew, term, trunc

    def step_servos_host_compact(self, action: np.ndarray):
        """Like ``step_servos_host`` but only the changing part of the observation crosses PCIe:
        returns ``obs[N, 6, 3]`` (position, velocity, torque per joint) and ``terminated``; temperature,
        voltage, reward and truncated are constants of the reference the caller fills once."""
        hb = self._host_buffers()
        a = action
        if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(action, dtype=np.float32)
        if a.size != self.n * 36:
            raise UpkieRuntimeError(f"action: expected {self.n * 36} float32 values, got {a.size}")
        obs, term = hb["obs18"], hb["term"]
        check(lib().upkie_b200_step_servos_host_compact(self._h, a.ctypes.data, obs.ctypes.data, term.ctypes.data))
        return obs, term

    def step_gyropod_host(self, action: np.ndarray):
        hb = self._host_buffers()
        a = np.ascontiguousarray(action, dtype=np.float32)
        act_dim = a.shape[1] if a.ndim == 2 else a.size // self.n
        if act_dim not in (1, 2) or a.size != self.n * act_dim:
            raise UpkieRuntimeError("action: expected shape [N, 1] (pendulum) or [N, 2] (gyropod)")
        obs = hb["obs6"] if act_dim == 2 else hb["obs4"]
        rew, term, trunc = hb["rew"], hb["term"], hb["trunc"]
        check(
            lib().upkie_b200_step_gyropod_host(self._h, a.ctypes.data, act_dim, obs.ctypes.data, None, term.ctypes.data, None)
        )
        return obs, rew, term, trunc

    @property
    def launches(self) -> int:
        """Step kernels launched through this handle, counted by the library (bench.py's ``gpu_launches``)."""
        c = C.c_uint64(0)
        check(lib().upkie_b200_launch_count(self._h, C.byref(c)))
        return int(c.value)

    # ------------------------------------------------------------------
    def spine_obs(self) -> torch.Tensor:
        """``get_spine_observation`` for all envs, flattened ``[N, 62]``."""
        out = torch.empty((self.n, _abi.SPINE_DIM), dtype=torch.float32, device=self.device)
        check(lib().upkie_b200_spine_obs(self._h, _ptr(out), self._stream()))
        return out

    def reset_obs(self, obs_dim: int) -> torch.Tensor:
        shape = (self.n, 6, 5) if obs_dim == 30 else (self.n, obs_dim)
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        check(lib().upkie_b200_reset_obs(self._h, int(obs_dim), _ptr(out), self._stream()))
        return out

    def get_state(self) -> torch.Tensor:
        out = torch.empty((self.n, _abi.STATE_DIM), dtype=torch.float32, device=self.device)
        check(lib().upkie_b200_get_state(self._h, _ptr(out), self._stream()))
        return out

    def get_body_contacts(self) -> torch.Tensor:
        """Body-ground contacts of the last substep, ``[n, BODY_REC_DIM]``: bit mask of the model's collision points that
        held contact rows, then per slot (collision point index, normal impulse, friction impulses along world -y and
        +x). What ``PyBulletBackend.get_contact_points`` reports for links other than the tires
        (``pybullet_backend.py:660-716``)."""
        out = torch.empty((self.n, _abi.BODY_REC_DIM), dtype=torch.float32, device=self.device)
        check(lib().upkie_b200_get_body_contacts(self._h, _ptr(out), self._stream()))
        return out

    def get_lag(self) -> torch.Tensor:
        """Spine mode: the lag records ``[N, LAG_DIM]`` (servo replies of the last two cycles, last IMU reading, the
        last assembled observation; ``include/upkie_b200.h`` ``UPKIE_LAG_*``)."""
        out = torch.empty((self.n, _abi.LAG_DIM), dtype=torch.float32, device=self.device)
        check(lib().upkie_b200_get_lag(self._h, _ptr(out), self._stream()))
        return out

    def set_lag(self, lag: torch.Tensor) -> None:
        self._check_tensor(lag, (self.n, _abi.LAG_DIM), name="lag")
        check(lib().upkie_b200_set_lag(self._h, _ptr(lag), self._stream()))

    def set_state(self, state: torch.Tensor) -> None:
        self._check_tensor(state, (self.n, _abi.STATE_DIM), name="state")
        check(lib().upkie_b200_set_state(self._h, _ptr(state), self._stream()))

    # checkpoint / resume ------------------------------------------------------------------
    def state_dict(self) -> dict:
        """Everything a handle needs to continue bit for bit (``torch.save``-able): robot state, episode / tick
        counters, pending auto-resets, error flags, randomisation, external forces, auto-reset keys. The model and
        the configuration are construction arguments and are not included."""
        i32, u8 = torch.int32, torch.uint8
        episode = torch.empty(self.n, dtype=i32, device=self.device)
        tick = torch.empty(self.n, dtype=i32, device=self.device)
        pending = torch.empty(self.n, dtype=u8, device=self.device)
        flags = torch.empty(self.n, dtype=i32, device=self.device)
        check(lib().upkie_b200_get_counters(self._h, _ptr(episode), _ptr(tick), _ptr(pending), _ptr(flags), self._stream()))
        friction, eps = getattr(self, "_randomization", (None, None))
        force, local_mask = getattr(self, "_external", (None, 0))
        return {
            "lag": self.get_lag() if self.config.spine_mode else None,  # spine mode: replies / IMU of the last cycles
            "state": self.get_state(), "episode": episode, "tick": tick, "pending_reset": pending, "error_flags": flags,
            "friction": friction, "inertia_eps": eps, "external_force": force, "external_local_mask": local_mask,
            "autoreset": getattr(self, "_autoreset", (AUTORESET_DISABLED, 0, 0)),
        }

    def load_state_dict(self, sd: dict) -> None:
        dev = self.device
        self.set_state(sd["state"].to(dev))
        if sd.get("lag") is not None:
            self.set_lag(sd["lag"].to(dev))
        check(lib().upkie_b200_set_counters(
            self._h, _ptr(sd["episode"].to(dev)), _ptr(sd["tick"].to(dev)), _ptr(sd["pending_reset"].to(dev)),
            _ptr(sd["error_flags"].to(dev)), self._stream()))
        torch.cuda.current_stream(dev).synchronize()
        self.set_randomization(None if sd["friction"] is None else sd["friction"].to(dev),
                               None if sd["inertia_eps"] is None else sd["inertia_eps"].to(dev))
        self.set_external_forces(None if sd["external_force"] is None else sd["external_force"].to(dev),
                                 sd["external_local_mask"])
        self.set_autoreset(*sd["autoreset"])

    def error_flags(self) -> torch.Tensor:
        out = torch.empty(self.n, dtype=torch.int32, device=self.device)
        check(lib().upkie_b200_error_flags(self._h, _ptr(out), self._stream()))
        return out


def neutral_action(model: Model, n: int, device=None) -> torch.Tensor:
    """``UpkieServos.get_neutral_action`` (``upkie_servos.py:255-262``) as a
    ``[N, 6, 6]`` tensor."""
    a = torch.zeros((n, 6, 6), dtype=torch.float32, device=device)
    a[:, :, 0] = float("nan")
    a[:, :, 3] = 1.0
    a[:, :, 4] = 1.0
    a[:, :, 5] = torch.tensor(model.tau_max, dtype=torch.float32, device=device)
    return a

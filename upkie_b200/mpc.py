# SPDX-License-Identifier: Apache-2.0
"""Batched model-predictive balancer.

``BatchedMPCBalancer`` is ``upkie.controllers.MPCBalancer``
(``upkie/controllers/mpc_balancer.py:127-312``) for N robots per launch: same
constructor parameters, same ``step`` semantics (fall detection, low-pass to
zero without floor contact, ``commanded_velocity += accel * dt / 2`` clamped to
``max_ground_velocity``), with the QP solved on the GPU (``csrc/mpc_core.cuh``).
"""

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _abi
from ._lib import check, lib
from .exceptions import UpkieRuntimeError


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class BatchedMPCBalancer:
    def __init__(
        self,
        n_robots: int,
        fall_pitch: float = 1.0,
        leg_length: float = 0.58,
        max_ground_accel: float = 10.0,
        max_ground_velocity: float = 3.0,
        nb_timesteps: int = 50,
        sampling_period: float = 0.02,
        stage_input_cost_weight: float = 1e-3,
        stage_state_cost_weight: float = 1e-3,
        terminal_cost_weight: float = 1.0,
        warm_start: bool = True,
        config: Optional[_abi.UpkieMpcConfig] = None,
        device: int = 0,
    ):
        if not torch.cuda.is_available():
            raise UpkieRuntimeError("upkie_b200 needs a CUDA device (there is no CPU fallback)")
        if config is None:
            config = _abi.default_mpc_config()
            config.fall_pitch = fall_pitch
            config.leg_length = leg_length
            config.max_ground_accel = max_ground_accel
            config.max_ground_velocity = max_ground_velocity
            config.nb_timesteps = nb_timesteps
            config.sampling_period = sampling_period
            config.stage_input_cost_weight = stage_input_cost_weight
            config.stage_state_cost_weight = stage_state_cost_weight
            config.terminal_cost_weight = terminal_cost_weight
        self.config = config
        self.n = int(n_robots)
        self.warm_start = warm_start
        self.device = torch.device("cuda", int(device))
        self.fall_pitch = config.fall_pitch
        self.max_ground_velocity = config.max_ground_velocity
        self._h = C.c_void_p()
        check(lib().upkie_b200_mpc_create(C.byref(config), self.n, int(device), C.byref(self._h)))
        #: ``MPCBalancer.commanded_velocity`` for every robot
        self.commanded_velocity = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.first_input = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.found = torch.zeros(self.n, dtype=torch.uint8, device=self.device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().upkie_b200_mpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, mask: Optional[torch.Tensor] = None) -> None:
        """``MPCBalancer.reset`` (``mpc_balancer.py:225-235``)."""
        if mask is None:
            self.commanded_velocity.zero_()
        else:
            self.commanded_velocity.masked_fill_(mask.bool(), 0.0)
        check(lib().upkie_b200_mpc_reset(self._h, _ptr(mask), self._stream()))

    def step_tensors(self, x0: torch.Tensor, v_target: torch.Tensor, floor_contact: Optional[torch.Tensor], dt: float):
        """x0[N, 4] = (ground position, pitch, ground velocity, pitch rate)."""
        if not self.warm_start:
            check(lib().upkie_b200_mpc_reset(self._h, None, self._stream()))
        check(
            lib().upkie_b200_mpc_step(
                self._h, _ptr(x0), _ptr(v_target), _ptr(floor_contact), float(dt), _ptr(self.commanded_velocity),
                _ptr(self.first_input), _ptr(self.found), self._stream(),
            )
        )
        return self.commanded_velocity

    def step_spine(self, target_ground_velocity: torch.Tensor, spine_obs: torch.Tensor, dt: float):
        """``MPCBalancer.step(target, spine_observation, dt)`` with the flat
        ``[N, 62]`` spine observation (``mpc_balancer.py:253-258``)."""
        from .base_velocity import mpc_inputs_from_spine

        x0, contact = mpc_inputs_from_spine(spine_obs)
        return self.step_tensors(x0, target_ground_velocity.contiguous(), contact, dt)

    def step(self, x0: np.ndarray, v_target: np.ndarray, floor_contact: np.ndarray, dt: float) -> np.ndarray:
        """Host-array convenience: H2D, solve, D2H."""
        xd = torch.from_numpy(np.ascontiguousarray(x0, dtype=np.float32)).to(self.device)
        vd = torch.from_numpy(np.ascontiguousarray(v_target, dtype=np.float32)).to(self.device)
        cd = torch.from_numpy(np.ascontiguousarray(floor_contact, dtype=np.uint8)).to(self.device)
        return self.step_tensors(xd, vd, cd, dt).cpu().numpy()

    def plan(self) -> torch.Tensor:
        """Optimal input sequence of the last solve, ``[N, nb_timesteps]``."""
        out = torch.empty((self.n, int(self.config.nb_timesteps)), dtype=torch.float32, device=self.device)
        check(lib().upkie_b200_mpc_plan(self._h, _ptr(out), self._stream()))
        return out

# SPDX-License-Identifier: Apache-2.0
"""Batched spine controller pipeline.

``WheelBalancerPipeline`` runs the reference spine's ``--pipeline wheel_balancer``
(``spines/common/controllers.h:24-44``) for N robots per launch: ``WheelStopper``
(``upkie/cpp/controllers/WheelStopper.cpp:15-22``) then the PI ``WheelBalancer``
(``upkie/cpp/controllers/WheelBalancer.cpp:35-110``). One ``step()`` = one
controller cycle of period ``1 / spine_frequency``.
"""

import ctypes as C
from typing import Optional

import torch

from . import _abi
from ._lib import check, lib
from .exceptions import UpkieRuntimeError


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class WheelBalancerPipeline:
    def __init__(self, n_robots: int, spine_frequency: float = 1000.0,
                 config: Optional[_abi.UpkieWheelBalancerConfig] = None, device: int = 0):
        if not torch.cuda.is_available():
            raise UpkieRuntimeError("upkie_b200 needs a CUDA device (there is no CPU fallback)")
        self.config = config if config is not None else _abi.default_wheel_balancer_config(spine_frequency)
        self.n = int(n_robots)
        self.device = torch.device("cuda", int(device))
        self._h = C.c_void_p()
        check(lib().upkie_b200_wheel_balancer_create(C.byref(self.config), self.n, int(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().upkie_b200_wheel_balancer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, mask: Optional[torch.Tensor] = None) -> None:
        """``WheelBalancer::reset`` (``WheelBalancer.cpp:26-33``): zero velocities, integrator and target."""
        check(lib().upkie_b200_wheel_balancer_reset(self._h, _ptr(mask), self._stream()))

    def step(self, obs: torch.Tensor, action: torch.Tensor, target: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``obs``: spine observation rows ``[N, 62]`` or observer rows ``[N, 21]`` (pitch, floor contact and
        wheel-odometry position are read from it); ``target[N, 2]``: target ground and yaw velocities (the
        ``"bullet"`` action key of the reference), default zeros; ``action[N, 6, 6]`` is updated in place and
        returned."""
        if obs.dtype != torch.float32 or not obs.is_contiguous() or obs.shape[0] != self.n:
            raise UpkieRuntimeError("obs: expected contiguous float32 [N, 62] or [N, 21]")
        if obs.shape[1] == _abi.SPINE_DIM:
            layout = _abi.OBS_LAYOUT_SPINE
        elif obs.shape[1] == _abi.OBSV_DIM:
            layout = _abi.OBS_LAYOUT_OBSERVERS
        else:
            raise UpkieRuntimeError("obs: expected spine rows [N, 62] or observer rows [N, 21]")
        if action.shape != (self.n, 6, 6) or action.dtype != torch.float32 or not action.is_contiguous():
            raise UpkieRuntimeError("action: expected contiguous float32 [N, 6, 6]")
        if target is not None and (target.shape != (self.n, 2) or target.dtype != torch.float32 or not target.is_contiguous()):
            raise UpkieRuntimeError("target: expected contiguous float32 [N, 2]")
        check(lib().upkie_b200_wheel_balancer_step(self._h, _ptr(obs), layout, _ptr(target), _ptr(action), self._stream()))
        return action

    def state(self) -> torch.Tensor:
        """``[N, 4]``: ground_velocity, integral_velocity, target_ground_position, target_yaw_velocity."""
        out = torch.empty((self.n, 4), dtype=torch.float32, device=self.device)
        check(lib().upkie_b200_wheel_balancer_state(self._h, _ptr(out), self._stream()))
        return out

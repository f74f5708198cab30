# SPDX-License-Identifier: Apache-2.0
"""Batched spine observer pipeline.

``ObserverPipeline`` runs the reference spine's observers -- ``BaseOrientation``
-> ``FloorContact`` (with one ``WheelContact`` per wheel) -> ``WheelOdometry``
(``spines/common/observers.h:23-44``, ``upkie/cpp/observers/*``) -- for N robots
per launch. One ``step()`` = one spine cycle (``dt = 1 / spine_frequency``).
"""

import ctypes as C
from typing import Optional

import torch

from . import _abi
from ._lib import check, lib
from .exceptions import UpkieRuntimeError
from .model import Model, default_model


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class ObserverPipeline:
    def __init__(self, n_robots: int, model: Optional[Model] = None, spine_frequency: float = 1000.0,
                 config: Optional[_abi.UpkieObserverConfig] = None, device: int = 0):
        if not torch.cuda.is_available():
            raise UpkieRuntimeError("upkie_b200 needs a CUDA device (there is no CPU fallback)")
        self.model = model if model is not None else default_model()
        self.config = config if config is not None else _abi.default_observer_config(self.model, spine_frequency)
        self.n = int(n_robots)
        self.device = torch.device("cuda", int(device))
        self._h = C.c_void_p()
        check(lib().upkie_b200_observers_create(C.byref(self.config), self.n, int(device), C.byref(self._h)))
        self.out = torch.zeros((self.n, _abi.OBSV_DIM), dtype=torch.float32, device=self.device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().upkie_b200_observers_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, mask: Optional[torch.Tensor] = None) -> None:
        """``Observer::reset`` of every observer (zero filters, odometry at 0)."""
        check(lib().upkie_b200_observers_reset(self._h, _ptr(mask), self._stream()))

    def step(self, spine_obs: torch.Tensor) -> torch.Tensor:
        """One cycle on the flat spine observation ``[N, 62]``; returns ``[N, 21]`` (``OBSV_*`` layout)."""
        if spine_obs.shape != (self.n, _abi.SPINE_DIM) or spine_obs.dtype != torch.float32 or not spine_obs.is_contiguous():
            raise UpkieRuntimeError("spine_obs: expected contiguous float32 [N, 62]")
        check(lib().upkie_b200_observers_step(self._h, _ptr(spine_obs), _ptr(self.out), self._stream()))
        return self.out

    @staticmethod
    def row_to_dict(r) -> dict:
        """Output row -> the observation sub-dictionaries the spine writes."""
        A = _abi
        return {
            "base_orientation": {
                "pitch": float(r[A.OBSV_PITCH]),
                "angular_velocity": [float(x) for x in r[A.OBSV_ANGVEL:A.OBSV_ANGVEL + 3]],
                "rotation_base_to_world": [[float(r[A.OBSV_ROT + 3 * i + j]) for j in range(3)] for i in range(3)],
            },
            "floor_contact": {
                "contact": bool(r[A.OBSV_CONTACT] > 0.5),
                "upper_leg_torque": float(r[A.OBSV_LEG_TORQUE]),
                "left_wheel": {"contact": bool(r[A.OBSV_WHEEL_CONTACT] > 0.5), "inertia": float(r[A.OBSV_WHEEL_INERTIA])},
                "right_wheel": {"contact": bool(r[A.OBSV_WHEEL_CONTACT + 1] > 0.5), "inertia": float(r[A.OBSV_WHEEL_INERTIA + 1])},
            },
            "wheel_odometry": {"position": float(r[A.OBSV_ODOM_POS]), "velocity": float(r[A.OBSV_ODOM_VEL])},
        }

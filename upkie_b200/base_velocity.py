# SPDX-License-Identifier: Apache-2.0
"""``UpkieBaseVelocity`` glue (``upkie/envs/upkie_base_velocity.py:164-202``) as device-agnostic tensor code.

The env is an MPC balancer in front of the gyropod env plus dead reckoning. The three heavy pieces (MPC solve,
gyropod step, spine observation) are passed in as callables: ``B200VectorEnv`` binds the CUDA kernels, the CPU tests
bind the oracle -- the SAME ordering logic runs in both, and is pinned on golden runs of the reference's own class
(``tests/test_base_velocity_golden.py``).
"""
from typing import Callable, Tuple

import torch

from . import _abi


def mpc_inputs_from_spine(spine_obs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``MPCBalancer.step`` observation unpacking (``mpc_balancer.py:253-258``): state
    ``[ground position, base pitch, ground velocity, base pitch rate]`` and the floor-contact flag from flat spine
    observation rows ``[N, 62]``."""
    A = _abi
    x0 = torch.stack(
        [
            spine_obs[:, A.SP_ODOM_POS],
            spine_obs[:, A.SP_PITCH],
            spine_obs[:, A.SP_ODOM_VEL],
            spine_obs[:, A.SP_BASE_ANGVEL + 1],
        ],
        dim=1,
    ).contiguous()
    contact = (spine_obs[:, A.SP_CONTACT] > 0.5).to(torch.uint8)
    return x0, contact


def base_velocity_tick(
    action: torch.Tensor,
    spine: torch.Tensor,
    xy: torch.Tensor,
    dt: float,
    mpc_step_spine: Callable[[torch.Tensor, torch.Tensor, float], torch.Tensor],
    step_gyropod: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]],
    spine_obs: Callable[[], torch.Tensor],
):
    """One ``UpkieBaseVelocity.step``: the MPC turns the commanded linear velocity ``action[:, 0]`` into a ground
    velocity from the LAST spine observation (the one the previous step or the reset returned), the gyropod env is
    stepped with ``[ground velocity, action[:, 1]]``, and ``xy`` (updated in place) dead-reckons the COMMANDED linear
    velocity along the POST-step yaw. Returns ``(obs[N, 3], reward, terminated, truncated, new_spine)``."""
    linear_velocity = action[:, 0].contiguous()
    ground_velocity = mpc_step_spine(linear_velocity, spine, dt)
    gyro_action = torch.stack([ground_velocity, action[:, 1]], dim=1).contiguous()
    obs6, rew, term, trunc = step_gyropod(gyro_action)
    new_spine = spine_obs()
    yaw = obs6[:, 2]
    xy[:, 0] += linear_velocity * torch.cos(yaw) * dt
    xy[:, 1] += linear_velocity * torch.sin(yaw) * dt
    obs = torch.cat([xy, yaw[:, None]], dim=1)
    return obs, rew, term, trunc, new_spine

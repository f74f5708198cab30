# SPDX-License-Identifier: Apache-2.0
"""upkie_b200 -- B200-native vectorised Upkie simulation and balance control.

Drop-in for the reference's env-step hot path (``UpkieServos`` /
``UpkieGyropod`` / ``UpkiePendulum`` / ``UpkieBaseVelocity`` on the PyBullet
backend, and ``MPCBalancer``); see DESIGN.md and INTEGRATION.md. Importing the
package needs neither a GPU nor the compiled library; creating an env does.
"""

__version__ = "0.1.0"

from . import _abi  # noqa: F401
from .exceptions import (  # noqa: F401
    MissingOptionalDependency,
    ModelError,
    UpkieException,
    UpkieRuntimeError,
)
from .model import ExternalForce, JointProperties, Model, default_model  # noqa: F401
from .robot_state import RobotState, RobotStateRandomization  # noqa: F401

ENV_IDS = {
    "Upkie-B200-Servos": "servos",
    "Upkie-B200-Gyropod": "gyropod",
    "Upkie-B200-Pendulum": "pendulum",
    "Upkie-B200-BaseVelocity": "base_velocity",
}


def make_vec(env_id: str, num_envs: int, **kwargs):
    """``gymnasium.make_vec``-style factory for the ids of ``register()``."""
    from .envs import B200VectorEnv

    return B200VectorEnv(num_envs, ENV_IDS[env_id], **kwargs)


def register() -> None:
    """Register ``Upkie-B200-{Servos,Gyropod,Pendulum,BaseVelocity}`` with Gymnasium,
    following the reference's ``Upkie-<Backend>-<Action>`` scheme
    (``upkie/envs/__init__.py:24-44``). Needs gymnasium."""
    try:
        import gymnasium as gym
    except ImportError as e:
        raise MissingOptionalDependency("gymnasium not found; use upkie_b200.make_vec(...) directly") from e
    for env_id, env_type in ENV_IDS.items():
        if env_id in gym.registry:
            continue

        def vector_entry_point(num_envs=1, _t=env_type, **kwargs):
            from .envs import B200VectorEnv

            return B200VectorEnv(num_envs, _t, **kwargs)

        gym.register(id=env_id, vector_entry_point=vector_entry_point)

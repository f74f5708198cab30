# SPDX-License-Identifier: Apache-2.0
"""Vectorised Upkie environments on the sm_100a kernels.

``B200VectorEnv`` is a ``gymnasium.vector.VectorEnv`` whose
``single_action_space`` / ``single_observation_space`` are identical to the
reference's single-robot environments:

- ``"servos"``   = ``UpkieServos``   (``upkie/envs/upkie_servos.py:20-344``)
- ``"gyropod"``  = ``UpkieGyropod``  (``upkie/envs/upkie_gyropod.py:20-392``)
- ``"pendulum"`` = ``UpkiePendulum`` (``upkie/envs/upkie_pendulum.py:20-142``)

with the reference's reset semantics (seeded NumPy sampling of the initial
state, ``upkie/envs/upkie_env.py:162-194``), ``reward == 0.0`` and
``truncated == False`` (``upkie_env.py:230-232``), fall termination for the
wheeled-inverted-pendulum wrappers (``upkie_gyropod.py:333-352``).
"""

from typing import Any, Dict, Optional, Tuple, Union

import numpy as np
import torch

from . import _abi
from .base_velocity import base_velocity_tick
from .exceptions import UpkieException, UpkieRuntimeError
from .gym_compat import VectorEnv, batch_space, spaces
from .model import Model, default_model
from .robot_state import RobotState
from .sim import AUTORESET_DISABLED, AUTORESET_NEXT_STEP, AUTORESET_SAME_STEP, UpkieSim

ENV_TYPES = ("servos", "gyropod", "pendulum", "base_velocity")
_AUTORESET = {"disabled": AUTORESET_DISABLED, "next_step": AUTORESET_NEXT_STEP, "same_step": AUTORESET_SAME_STEP}


# ---- spaces (host-side, no GPU needed) ------------------------------------------------

def make_servo_spaces(model: Model, max_gain_scale: float = 5.0):
    """``UpkieServos.__create_servo_spaces`` (``upkie_servos.py:173-286``).

    Returns ``(action_space, observation_space, neutral_action, max_action,
    min_action)``.
    """
    if not (0.0 < max_gain_scale < 10.0):
        raise UpkieRuntimeError(f"Invalid value {max_gain_scale=}")
    action_space, servo_space = {}, {}
    neutral_action, max_action, min_action = {}, {}, {}
    f32 = np.float32

    def box(lo, hi):
        return spaces.Box(low=lo, high=hi, shape=(1,), dtype=f32)

    for joint in model.joints:
        lim = joint.limit
        action_space[joint.name] = spaces.Dict(
            {
                "position": box(lim.lower, lim.upper),
                "velocity": box(-lim.velocity, +lim.velocity),
                "feedforward_torque": box(-lim.effort, +lim.effort),
                "kp_scale": box(0.0, max_gain_scale),
                "kd_scale": box(0.0, max_gain_scale),
                "maximum_torque": box(0.0, lim.effort),
            }
        )
        servo_space[joint.name] = spaces.Dict(
            {
                "position": box(lim.lower, lim.upper),
                "velocity": box(-lim.velocity, +lim.velocity),
                "torque": box(-lim.effort, +lim.effort),
                "temperature": box(0.0, 100.0),
                "voltage": box(10.0, 44.0),  # moteus min 10 V, max 44 V
            }
        )
        neutral_action[joint.name] = {
            "position": np.nan,
            "velocity": 0.0,
            "feedforward_torque": 0.0,
            "kp_scale": 1.0,
            "kd_scale": 1.0,
            "maximum_torque": lim.effort,
        }
        max_action[joint.name] = {
            "position": lim.upper,
            "velocity": lim.velocity,
            "feedforward_torque": lim.effort,
            "kp_scale": max_gain_scale,
            "kd_scale": max_gain_scale,
            "maximum_torque": lim.effort,
        }
        min_action[joint.name] = {
            "position": lim.lower,
            "velocity": -lim.velocity,
            "feedforward_torque": -lim.effort,
            "kp_scale": 0.0,
            "kd_scale": 0.0,
            "maximum_torque": 0.0,
        }
    return spaces.Dict(action_space), spaces.Dict(servo_space), neutral_action, max_action, min_action


def make_gyropod_spaces(max_ground_velocity: float = 3.0, max_yaw_velocity: float = 1.0):
    """``UpkieGyropod.__init__`` spaces (``upkie_gyropod.py:118-160``)."""
    observation_limit = np.array(
        [float("inf"), np.pi, float("inf"), max_ground_velocity, 1000.0, max_yaw_velocity], dtype=np.float32
    )
    action_limit = np.array([max_ground_velocity, max_yaw_velocity], dtype=np.float32)
    obs = spaces.Box(-observation_limit, +observation_limit, shape=observation_limit.shape, dtype=observation_limit.dtype)
    act = spaces.Box(-action_limit, +action_limit, shape=action_limit.shape, dtype=action_limit.dtype)
    return act, obs


PENDULUM_OBS_INDICES = [1, 0, 4, 3]  # upkie_pendulum.py:17


def make_pendulum_spaces(max_ground_velocity: float = 3.0):
    """``UpkiePendulum.__init__`` spaces (``upkie_pendulum.py:87-102``)."""
    _, gyro_obs = make_gyropod_spaces(max_ground_velocity)
    obs_limit = gyro_obs.high[PENDULUM_OBS_INDICES]
    obs = spaces.Box(-obs_limit, +obs_limit, shape=obs_limit.shape, dtype=np.float32)
    action_limit = np.array([max_ground_velocity], dtype=np.float32)
    act = spaces.Box(-action_limit, +action_limit, shape=action_limit.shape, dtype=np.float32)
    return act, obs


def make_base_velocity_spaces(max_ground_velocity: float = 3.0, max_yaw_velocity: float = 1.0):
    """``UpkieBaseVelocity.__init__`` spaces (``upkie_base_velocity.py:96-115``)."""
    observation_limit = np.full(3, float("inf"), dtype=np.float32)
    action_limit = np.array([max_ground_velocity, max_yaw_velocity], dtype=np.float32)
    obs = spaces.Box(-observation_limit, +observation_limit, shape=observation_limit.shape, dtype=observation_limit.dtype)
    act = spaces.Box(-action_limit, +action_limit, shape=action_limit.shape, dtype=action_limit.dtype)
    return act, obs


def servo_action_dict_to_array(action: dict, neutral_action: dict, n: int) -> np.ndarray:
    """Batched dict action ``{joint: {key: array[N, 1]}}`` -> ``[N, 6, 6]``.
    Missing keys take the neutral action (``upkie_servos.py:326-331``)."""
    out = np.empty((n, 6, 6), dtype=np.float32)
    for j, name in enumerate(_abi.JOINT_NAMES):
        ja = action.get(name, {}) if isinstance(action, dict) else {}
        for k, key in enumerate(_abi.ACT_KEYS):
            if key in ja:
                out[:, j, k] = np.asarray(ja[key], dtype=np.float32).reshape(n)
            else:
                out[:, j, k] = neutral_action[name][key]
    return out


def servo_obs_array_to_dict(obs: np.ndarray) -> dict:
    """``[N, 6, 5]`` -> batched dict ``{joint: {key: array[N, 1] float32}}``
    (``UpkieServos.get_env_observation``, ``upkie_servos.py:288-306``)."""
    return {
        name: {key: obs[:, j, k : k + 1].astype(np.float32, copy=False) for k, key in enumerate(_abi.OBS_KEYS)}
        for j, name in enumerate(_abi.JOINT_NAMES)
    }


class SpineObservations:
    """Lazy ``info["spine_observation"]``: fetched from the device on first use.

    ``obs[i]`` returns the reference's nested dictionary for env ``i``
    (``pybullet_backend.py:325-331``); ``obs.array`` the flat ``[N, 62]`` array.
    """

    def __init__(self, sim: UpkieSim):
        self._sim = sim
        self._array = None
        self._stamp = sim.launches  # step kernels launched so far: identifies the tick this object belongs to

    @property
    def array(self) -> np.ndarray:
        if self._array is None:
            if self._sim.launches != self._stamp:
                # the simulator has moved on: fetching now would silently return a LATER tick's observation
                raise UpkieRuntimeError(
                    "info['spine_observation'] is fetched lazily and the simulator has been stepped since this step: "
                    "read it (e.g. `.array`) before calling step() again")
            self._array = self._sim.spine_obs().cpu().numpy()
        return self._array

    def __len__(self):
        return self._sim.n

    def __getitem__(self, i: int) -> dict:
        return spine_row_to_dict(self.array[i])


def spine_row_to_dict(r: np.ndarray) -> dict:
    """Flat spine observation row -> the dictionary of
    ``PyBulletBackend.get_spine_observation``."""
    A = _abi
    return {
        "base_orientation": {
            "angular_velocity": [float(x) for x in r[A.SP_BASE_ANGVEL : A.SP_BASE_ANGVEL + 3]],
            "linear_velocity": [float(x) for x in r[A.SP_BASE_LINVEL : A.SP_BASE_LINVEL + 3]],
            "pitch": float(r[A.SP_PITCH]),
            "rotation_base_to_world": np.asarray(r[A.SP_ROT : A.SP_ROT + 9], dtype=float).reshape(3, 3).tolist(),
        },
        "floor_contact": {"contact": bool(r[A.SP_CONTACT] > 0.5)},
        "imu": {
            "orientation": [float(x) for x in r[A.SP_IMU_QUAT : A.SP_IMU_QUAT + 4]],
            "angular_velocity": [float(x) for x in r[A.SP_IMU_ANGVEL : A.SP_IMU_ANGVEL + 3]],
            "linear_acceleration": np.asarray(r[A.SP_IMU_LINACC : A.SP_IMU_LINACC + 3], dtype=float),
            "raw_linear_acceleration": [float(x) for x in r[A.SP_IMU_RAWACC : A.SP_IMU_RAWACC + 3]],
        },
        "servo": {
            name: {
                key: float(r[A.SP_SERVO + j * 5 + k]) for k, key in enumerate(A.OBS_KEYS)
            }
            for j, name in enumerate(A.JOINT_NAMES)
        },
        "wheel_odometry": {"position": float(r[A.SP_ODOM_POS]), "velocity": float(r[A.SP_ODOM_VEL])},
    }


def make_config(
    frequency: float = 200.0,
    nb_substeps: Optional[int] = None,
    torque_control_kp: float = 20.0,
    torque_control_kd: float = 1.0,
    joint_properties: Optional[dict] = None,
    max_gain_scale: float = 5.0,
    fall_pitch: float = 1.0,
    leg_gain_scale: float = 1.0,
    max_ground_velocity: float = 3.0,
    max_yaw_velocity: float = 1.0,
    init_state: Optional[RobotState] = None,
    noise_seed: int = 0,
    joint_limits: Union[bool, int] = True,
    spine_mode: bool = False,
    body_contacts: bool = False,
) -> _abi.UpkieSimConfig:
    """Split of the keyword arguments the reference's factories forward to the
    backend, the servo env and the wrappers (``upkie/envs/entry_points.py:41-61,99-109``)."""
    if frequency is None:
        raise UpkieException("This environment needs a loop frequency")
    cfg = _abi.default_sim_config(frequency)
    if nb_substeps is not None:
        cfg.nb_substeps = int(nb_substeps)
    if cfg.nb_substeps < 1:
        cfg.nb_substeps = 1
    cfg.torque_control_kp = torque_control_kp
    cfg.torque_control_kd = torque_control_kd
    for j, name in enumerate(_abi.JOINT_NAMES):
        props = (joint_properties or {}).get(name)
        if props is not None:
            cfg.joint_friction[j] = float(getattr(props, "friction", 0.0))
            cfg.torque_control_noise[j] = float(getattr(props, "torque_control_noise", 0.0))
            cfg.torque_measurement_noise[j] = float(getattr(props, "torque_measurement_noise", 0.0))
    cfg.noise_seed = int(noise_seed) & 0xFFFFFFFFFFFFFFFF
    # Bullet's joint-limit constraint rows on hips and knees (include/upkie_b200.h: joint_limits): on, as in the
    # multibody PyBullet's importer builds (pybullet_backend.py:121). True -> 3 (the packed ten-row solver for the warps
    # that hold a robot on a bound); 2 = ten-row solver for every warp; False / 0 = no limit rows (round-1 behaviour)
    cfg.joint_limits = 3 if joint_limits is True else int(joint_limits)
    # timing of the C++ Bullet spine in simulate() mode instead of PyBulletBackend's (include/upkie_b200.h: spine_mode)
    cfg.spine_mode = 1 if spine_mode else 0
    # body-ground contacts (include/upkie_b200.h: body_contacts): the collision shapes of the model's links other than
    # the tires hold contact rows against the floor, as they do in Bullet; needs the joint-limit kernels
    cfg.body_contacts = 1 if (body_contacts and cfg.joint_limits) else 0
    cfg.max_gain_scale = max_gain_scale
    cfg.fall_pitch = fall_pitch
    cfg.leg_gain_scale = leg_gain_scale
    cfg.max_ground_velocity = max_ground_velocity
    cfg.max_yaw_velocity = max_yaw_velocity
    if init_state is not None:
        init_state.apply_to_config(cfg)
    return cfg


class B200VectorEnv(VectorEnv):
    """N Upkie environments stepped by one kernel launch per ``step()``."""

    metadata: Dict[str, Any] = {"autoreset_mode": "disabled"}

    def __init__(
        self,
        num_envs: int,
        env_type: str = "servos",
        frequency: float = 200.0,
        init_state: Optional[RobotState] = None,
        model: Optional[Model] = None,
        device: int = 0,
        autoreset_mode: str = "disabled",
        max_gain_scale: float = 5.0,
        fall_pitch: float = 1.0,
        leg_gain_scale: float = 1.0,
        max_ground_velocity: float = 3.0,
        max_yaw_velocity: float = 1.0,
        nb_substeps: Optional[int] = None,
        torque_control_kp: float = 20.0,
        torque_control_kd: float = 1.0,
        joint_properties: Optional[dict] = None,
        inertia_variation: float = 0.0,
        env_offset: int = 0,
        config: Optional[_abi.UpkieSimConfig] = None,
        leg_length: float = 0.58,
        max_ground_accel: float = 10.0,
        noise_seed: int = 0,
        joint_limits: Union[bool, int] = True,
        copy: bool = True,
        spine_mode: bool = False,
        body_contacts: bool = False,
    ):
        if env_type not in ENV_TYPES:
            raise UpkieException(f"env_type must be one of {ENV_TYPES}")
        if autoreset_mode not in _AUTORESET:
            raise UpkieException(f"autoreset_mode must be one of {tuple(_AUTORESET)}")
        self.env_type = env_type
        # host path: ``step()`` results live in the handle's pinned output buffers, which the next ``step()``
        # overwrites. copy=True (default, as gymnasium.vector.SyncVectorEnv(copy=True)) hands out copies, so that
        # ``buf.append(obs)`` or ``prev_obs`` comparisons do not alias; copy=False returns views of the pinned buffers
        # (zero-copy fast path, valid until the next step; what bench.py's e2e figure uses and says so)
        self.copy = bool(copy)
        self.num_envs = int(num_envs)
        self.model = model if model is not None else default_model()
        self.init_state = init_state if init_state is not None else RobotState(
            position_base_in_world=np.array([0.0, 0.0, 0.6])
        )
        self.frequency = frequency
        self.dt = 1.0 / frequency
        self.env_offset = int(env_offset)
        self.autoreset_mode = autoreset_mode
        self.metadata = dict(self.metadata, autoreset_mode=autoreset_mode)
        if config is None:
            config = make_config(
                frequency, nb_substeps, torque_control_kp, torque_control_kd, joint_properties, max_gain_scale,
                fall_pitch, leg_gain_scale, max_ground_velocity, max_yaw_velocity, self.init_state, noise_seed,
                joint_limits, spine_mode, body_contacts,
            )
        self.config = config
        if self.config.spine_mode and env_type != "servos":
            raise UpkieException("spine_mode is available for env_type='servos' (the wrappers of a spine read observer "
                                 "outputs: feed info['spine_observation'] to upkie_b200.observers.ObserverPipeline)")
        (
            servo_act, servo_obs, self._neutral_action, self._max_action, self._min_action,
        ) = make_servo_spaces(self.model, max_gain_scale)
        if env_type == "servos":
            self.single_action_space, self.single_observation_space = servo_act, servo_obs
        elif env_type == "gyropod":
            self.single_action_space, self.single_observation_space = make_gyropod_spaces(
                max_ground_velocity, max_yaw_velocity
            )
        elif env_type == "base_velocity":
            if autoreset_mode != "disabled":
                raise UpkieException("base_velocity envs support autoreset_mode='disabled' only")
            self.single_action_space, self.single_observation_space = make_base_velocity_spaces(
                max_ground_velocity, max_yaw_velocity
            )
        else:
            self.single_action_space, self.single_observation_space = make_pendulum_spaces(max_ground_velocity)
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        self.observation_space = batch_space(self.single_observation_space, self.num_envs)

        self.sim = UpkieSim(self.num_envs, model=self.model, config=self.config, device=device)
        self._seed = 0
        self.mpc_balancer = None
        if env_type == "base_velocity":
            # UpkieBaseVelocity embeds an MPCBalancer (upkie_base_velocity.py:117-126)
            from .mpc import BatchedMPCBalancer

            self.mpc_balancer = BatchedMPCBalancer(
                self.num_envs, fall_pitch=fall_pitch, leg_length=leg_length, max_ground_accel=max_ground_accel,
                max_ground_velocity=max_ground_velocity, device=device,
            )
            self._xy = torch.zeros((self.num_envs, 2), dtype=torch.float32, device=self.sim.device)
            self._spine = None
        self.sim.set_autoreset(_AUTORESET[autoreset_mode], self._seed, self.env_offset)
        self.inertia_variation = inertia_variation
        if abs(inertia_variation) > 1e-10:
            self.randomize_inertias(inertia_variation)

    # ------------------------------------------------------------------
    def get_neutral_action(self) -> dict:
        """``UpkieServos.get_neutral_action`` (``upkie_servos.py:308-314``)."""
        return self._neutral_action.copy()

    def randomize_inertias(self, inertia_variation: float, seed: Optional[int] = None) -> None:
        """``PyBulletBackend.randomize_inertias`` (``pybullet_backend.py:571-601``):
        one epsilon ~ U(-v, v) per non-base body and env."""
        rng = np.random.default_rng(seed)
        eps = rng.uniform(-inertia_variation, inertia_variation, size=(self.num_envs, 6)).astype(np.float32)
        self.sim.set_randomization(inertia_eps=torch.from_numpy(eps).to(self.sim.device))

    def update_init_rand(self, **kwargs) -> None:
        """``UpkieEnv.update_init_rand`` (``upkie_env.py:244-251``)."""
        self.init_state.randomization.update(**kwargs)
        # the fused auto-reset samples on the device: carry the new bounds there too
        self.init_state.apply_to_config(self.config)
        self.sim.set_config(self.config)

    def close(self, **kwargs) -> None:
        self.sim.close()

    # ------------------------------------------------------------------
    def _obs_dim(self) -> int:
        return {"servos": 30, "gyropod": 6, "pendulum": 4, "base_velocity": 6}[self.env_type]

    def _format_obs(self, obs: np.ndarray):
        if self.env_type != "servos":
            return obs
        # the host-path observation lives in one persistent pinned buffer: build the dictionary of
        # views once and hand the same (in-place updated) dictionary back on every step
        cache = getattr(self, "_obs_dict_cache", None)
        if cache is None or cache[0] is not obs:
            cache = (obs, servo_obs_array_to_dict(obs))
            if obs is self.sim._host_buffers()["obs30"]:
                self._obs_dict_cache = cache
        return cache[1]

    def get_contact_points(self, env_index: int = 0, link_name: Optional[str] = None) -> list:
        """``PyBulletBackend.get_contact_points`` (``pybullet_backend.py:660-716``) for one env of the batch."""
        from .model import contact_points_from_state

        row = self.sim.get_state()[int(env_index)].cpu().numpy()
        rec = self.sim.get_body_contacts()[int(env_index)].cpu().numpy()
        return contact_points_from_state(self.model, row, self.config, link_name, rec)

    def set_external_forces(self, external_forces: Optional[dict]) -> None:
        """Batched ``PyBulletBackend.set_external_forces``: ``{link name: ExternalForce}`` whose ``force`` is
        ``[3]`` (all envs) or ``[N, 3]``; ``None`` or ``{}`` clears."""
        if not external_forces:
            self.sim.set_external_forces(None)
            return
        rows, mask = self.model.external_force_rows(external_forces, self.num_envs)
        self.sim.set_external_forces(torch.from_numpy(rows).to(self.sim.device), mask)

    def _servo_obs_dict(self, obs18: np.ndarray, cache: bool = True) -> dict:
        """Batched observation dictionary over the persistent pinned ``[N, 6, 3]`` buffer the kernel writes
        (position, velocity, torque): built once, updated in place by every step. Temperature and voltage
        are the simulator's constants (``pybullet_backend.py:471-472``) and never cross PCIe."""
        cached = getattr(self, "_obs18_cache", None) if cache else None
        if cached is None or cached[0] is not obs18:
            n = self.num_envs
            const = getattr(self, "_obs_constants", None)
            if const is None:
                # pybullet_backend.py:471 (42.0) / BulletInterface.cpp:70 in spine mode (20.0)
                temperature = np.full((n, 1), 20.0 if self.config.spine_mode else 42.0, dtype=np.float32)
                voltage = np.full((n, 1), 18.0, dtype=np.float32)
                temperature.flags.writeable = False
                voltage.flags.writeable = False
                const = self._obs_constants = (temperature, voltage)
            temperature, voltage = const
            d = {
                name: {
                    "position": obs18[:, j, 0:1],
                    "velocity": obs18[:, j, 1:2],
                    "torque": obs18[:, j, 2:3],
                    "temperature": temperature,
                    "voltage": voltage,
                }
                for j, name in enumerate(_abi.JOINT_NAMES)
            }
            cached = (obs18, d)
            if cache:
                self._obs18_cache = cached
        return cached[1]

    def reset(self, *, seed: Optional[Union[int, list]] = None, options: Optional[dict] = None):
        """Reset all envs (or ``options["reset_mask"]``) and return the initial
        observations. Env ``i`` samples its initial state from
        ``np.random.default_rng(seed + i)`` exactly as ``UpkieEnv.reset(seed)``
        does for a single env (``upkie_env.py:180-190``); with ``seed=None``
        the envs draw from the vector env's own generator."""
        n = self.num_envs
        mask = None
        if options and options.get("reset_mask") is not None:
            mask = np.ascontiguousarray(options["reset_mask"], dtype=np.uint8).reshape(n)
        parent = self.np_random
        if seed is None:
            seeds = [None] * n
        elif isinstance(seed, (list, tuple, np.ndarray)):
            seeds = list(seed)
        else:
            seeds = [int(seed) + self.env_offset + i for i in range(n)]
            self._seed = int(seed)
            self.sim.set_autoreset(_AUTORESET[self.autoreset_mode], self._seed, self.env_offset)
        rows = np.zeros((n, _abi.INIT_DIM), dtype=np.float32)
        for i in range(n):
            if mask is not None and not mask[i]:
                continue
            rng = parent if seeds[i] is None else np.random.default_rng(seeds[i])
            rows[i] = self.init_state.sample_state(rng).to_row()
        dev = self.sim.device
        self.sim.reset(
            mask=torch.from_numpy(mask).to(dev) if mask is not None else None,
            init_state=torch.from_numpy(rows).to(dev),
        )
        info = {"spine_observation": SpineObservations(self.sim)}
        if self.env_type == "base_velocity":
            # UpkieBaseVelocity.reset (upkie_base_velocity.py:138-162): MPC reset, x = y = 0, zero observation
            self.mpc_balancer.reset()
            self._xy.zero_()
            self._spine = self.sim.spine_obs()
            return np.zeros((n, 3), dtype=np.float32), info
        obs = self.sim.reset_obs(self._obs_dim()).cpu().numpy()
        return self._format_obs(obs), info

    def step(self, action):
        """One 5 ms control tick for every env. ``action`` is a batched dict
        (servos), an ndarray / CPU tensor, or a CUDA tensor (then the result
        tensors stay on the device)."""
        if isinstance(action, torch.Tensor) and action.is_cuda:
            return self.step_tensors(action)
        n = self.num_envs
        if self.env_type == "base_velocity":
            a = torch.from_numpy(np.ascontiguousarray(np.asarray(action, dtype=np.float32).reshape(n, 2))).to(self.sim.device)
            obs, rew, term, trunc, info = self.step_tensors(a)
            return obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy().view(np.bool_), trunc.cpu().numpy().view(np.bool_), info
        if self.env_type == "servos":
            a = (
                servo_action_dict_to_array(action, self._neutral_action, n)
                if isinstance(action, dict)
                else np.ascontiguousarray(np.asarray(action, dtype=np.float32).reshape(n, 6, 6))
            )
            obs18, term = self.sim.step_servos_host_compact(a)
            hb = self.sim._host_buffers()
            info = {"spine_observation": SpineObservations(self.sim)}
            if self.copy:
                obs18 = obs18.copy()
                return (self._servo_obs_dict(obs18, cache=False), hb["rew"].copy(), term.view(np.bool_).copy(),
                        hb["trunc"].view(np.bool_).copy(), info)
            return self._servo_obs_dict(obs18), hb["rew"], term.view(np.bool_), hb["trunc"].view(np.bool_), info
        else:
            d = 2 if self.env_type == "gyropod" else 1
            a = np.ascontiguousarray(np.asarray(action, dtype=np.float32).reshape(n, d))
            obs, rew, term, trunc = self.sim.step_gyropod_host(a)
        info = {"spine_observation": SpineObservations(self.sim)}
        if self.copy:
            return obs.copy(), rew.copy(), term.view(np.bool_).copy(), trunc.view(np.bool_).copy(), info
        # views of the handle's pinned output buffers: valid until the next step()
        return self._format_obs(obs), rew, term.view(np.bool_), trunc.view(np.bool_), info

    def step_tensors(self, action: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, dict]:
        """Zero-copy fast path: CUDA tensors in, CUDA tensors out
        (``action[N, 6, 6]`` / ``[N, 2]`` / ``[N, 1]``)."""
        if self.env_type == "servos":
            obs, rew, term, trunc = self.sim.step_servos(action)
        elif self.env_type == "gyropod":
            obs, rew, term, trunc = self.sim.step_gyropod(action)
        elif self.env_type == "base_velocity":
            # UpkieBaseVelocity.step (upkie_base_velocity.py:164-202): the MPC turns the commanded linear
            # velocity into a ground velocity from the LAST spine observation, the gyropod env is stepped,
            # (x, y) dead-reckon the commanded velocity along the post-step yaw
            if self._spine is None:
                self._spine = self.sim.spine_obs()
            obs, rew, term, trunc, self._spine = base_velocity_tick(
                action, self._spine, self._xy, self.dt, self.mpc_balancer.step_spine, self.sim.step_gyropod,
                self.sim.spine_obs,
            )
        else:
            obs, rew, term, trunc = self.sim.step_pendulum(action)
        return obs, rew, term, trunc, {"spine_observation": SpineObservations(self.sim)}

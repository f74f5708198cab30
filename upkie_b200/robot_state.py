# SPDX-License-Identifier: Apache-2.0
"""Initial robot state with optional randomisation.

Same classes, fields and sampling order as the reference
(``upkie/utils/robot_state.py:14-196``,
``upkie/utils/robot_state_randomization.py:14-189``); orientations are stored as
``(w, x, y, z)`` quaternions, and ``scipy.spatial.transform.Rotation`` objects
are accepted wherever the reference takes one.
"""

from typing import Optional

import numpy as np

from . import _abi


def _quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.array(
        [
            a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
            a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
        ]
    )


def quat_from_euler_zyx(yaw: float, pitch: float, roll: float) -> np.ndarray:
    """Quaternion (w, x, y, z) of intrinsic ZYX euler angles, i.e. what
    ``ScipyRotation.from_euler("ZYX", [yaw, pitch, roll])`` represents."""
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    cp, sp = np.cos(pitch / 2), np.sin(pitch / 2)
    cr, sr = np.cos(roll / 2), np.sin(roll / 2)
    return np.array(
        [
            cy * cp * cr + sy * sp * sr,
            cy * cp * sr - sy * sp * cr,
            cy * sp * cr + sy * cp * sr,
            sy * cp * cr - cy * sp * sr,
        ]
    )


def _as_quat_wxyz(orientation) -> np.ndarray:
    if orientation is None:
        return np.array([1.0, 0.0, 0.0, 0.0])
    if hasattr(orientation, "as_quat"):  # scipy Rotation: (x, y, z, w)
        x, y, z, w = orientation.as_quat()
        return np.array([w, x, y, z], dtype=float)
    q = np.asarray(orientation, dtype=float).reshape(4)
    return q / np.linalg.norm(q)


def _rotation_from_quat_wxyz(q: np.ndarray):
    try:
        from scipy.spatial.transform import Rotation

        return Rotation.from_quat([q[1], q[2], q[3], q[0]])
    except ImportError:  # pragma: no cover
        return q


class RobotStateRandomization:
    """Domain randomisation of the initial state
    (``robot_state_randomization.py:14-189``)."""

    def __init__(
        self,
        roll: float = 0.0,
        pitch: float = 0.0,
        x: float = 0.0,
        z: float = 0.0,
        omega_x: float = 0.0,
        omega_y: float = 0.0,
        linear_velocity: Optional[np.ndarray] = None,
    ):
        self.roll = roll
        self.pitch = pitch
        self.x = x
        self.z = z
        self.omega_x = omega_x
        self.omega_y = omega_y
        self.linear_velocity = (
            np.asarray(linear_velocity, dtype=float) if linear_velocity is not None else np.zeros(3)
        )

    def update(self, roll=None, pitch=None, x=None, z=None, omega_x=None, omega_y=None, v_x=None, v_z=None) -> None:
        if roll is not None:
            self.roll = roll
        if pitch is not None:
            self.pitch = pitch
        if x is not None:
            self.x = x
        if z is not None:
            self.z = z
        if omega_x is not None:
            self.omega_x = omega_x
        if omega_y is not None:
            self.omega_y = omega_y
        if v_x is not None:
            self.linear_velocity[0] = v_x
        if v_z is not None:
            self.linear_velocity[2] = v_z

    def sample_orientation_quat(self, np_random: np.random.Generator) -> np.ndarray:
        """Quaternion (w, x, y, z) of the sampled rotation (``:135-150``)."""
        yaw_pitch_roll_bounds = np.array([0.0, self.pitch, self.roll])
        ypr = np_random.uniform(low=-yaw_pitch_roll_bounds, high=+yaw_pitch_roll_bounds, size=3)
        return quat_from_euler_zyx(ypr[0], ypr[1], ypr[2])

    def sample_orientation(self, np_random: np.random.Generator):
        """As in the reference: a scipy ``Rotation`` (the (w, x, y, z) quaternion when scipy is missing)."""
        return _rotation_from_quat_wxyz(self.sample_orientation_quat(np_random))

    def sample_position(self, np_random: np.random.Generator) -> np.ndarray:
        return np_random.uniform(
            low=np.array([-self.x, 0.0, 0.0]), high=np.array([+self.x, 0.0, self.z]), size=3
        )

    def sample_angular_velocity(self, np_random: np.random.Generator) -> np.ndarray:
        return np_random.uniform(
            low=np.array([-self.omega_x, -self.omega_y, 0.0]),
            high=np.array([+self.omega_x, +self.omega_y, 0.0]),
            size=3,
        )

    def sample_linear_velocity(self, np_random: np.random.Generator) -> np.ndarray:
        return np_random.uniform(low=-self.linear_velocity, high=self.linear_velocity, size=3)


class RobotState:
    """Robot state (configuration and velocity) with optional randomisation
    (``robot_state.py:14-196``)."""

    def __init__(
        self,
        angular_velocity_base_in_base: Optional[np.ndarray] = None,
        joint_configuration: Optional[np.ndarray] = None,
        joint_velocity: Optional[np.ndarray] = None,
        linear_velocity_base_to_world_in_world: Optional[np.ndarray] = None,
        orientation_base_in_world=None,
        position_base_in_world: Optional[np.ndarray] = None,
        randomization: Optional[RobotStateRandomization] = None,
    ):
        self.angular_velocity_base_in_base = (
            np.asarray(angular_velocity_base_in_base, dtype=float)
            if angular_velocity_base_in_base is not None
            else np.zeros(3)
        )
        self.joint_configuration = (
            np.asarray(joint_configuration, dtype=float) if joint_configuration is not None else np.zeros(6)
        )
        self.joint_velocity = np.asarray(joint_velocity, dtype=float) if joint_velocity is not None else np.zeros(6)
        self.linear_velocity_base_to_world_in_world = (
            np.asarray(linear_velocity_base_to_world_in_world, dtype=float)
            if linear_velocity_base_to_world_in_world is not None
            else np.zeros(3)
        )
        self.orientation_quat_wxyz = _as_quat_wxyz(orientation_base_in_world)
        self.position_base_in_world = (
            np.asarray(position_base_in_world, dtype=float)
            if position_base_in_world is not None
            else np.array([0.0, 0.0, 0.6])  # Upkie above horizontal plane
        )
        self.randomization = randomization if randomization is not None else RobotStateRandomization()

    @property
    def orientation_base_in_world(self):
        """scipy ``Rotation`` (as in the reference) when scipy is importable,
        else the (w, x, y, z) quaternion."""
        try:
            from scipy.spatial.transform import Rotation

            w, x, y, z = self.orientation_quat_wxyz
            return Rotation.from_quat([x, y, z, w])
        except ImportError:  # pragma: no cover
            return self.orientation_quat_wxyz

    def sample_angular_velocity(self, np_random):
        return self.angular_velocity_base_in_base + self.randomization.sample_angular_velocity(np_random)

    def sample_linear_velocity(self, np_random):
        return self.linear_velocity_base_to_world_in_world + self.randomization.sample_linear_velocity(np_random)

    def sample_orientation_quat(self, np_random) -> np.ndarray:
        # rotation_base_to_world * rotation_rand_to_base
        return _quat_mul(self.orientation_quat_wxyz, self.randomization.sample_orientation_quat(np_random))

    def sample_orientation(self, np_random):
        """As in the reference (``robot_state.py:150-163``): a scipy ``Rotation``."""
        return _rotation_from_quat_wxyz(self.sample_orientation_quat(np_random))

    def sample_position(self, np_random):
        return self.position_base_in_world + self.randomization.sample_position(np_random)

    def sample_state(self, np_random: np.random.Generator) -> "RobotState":
        """Draw order of the reference (``:182-187``): angular velocity, linear
        velocity, orientation, position."""
        sampled_angular_velocity = self.sample_angular_velocity(np_random)
        sampled_linear_velocity = self.sample_linear_velocity(np_random)
        sampled_orientation = self.sample_orientation_quat(np_random)
        sampled_position = self.sample_position(np_random)
        return RobotState(
            angular_velocity_base_in_base=sampled_angular_velocity,
            joint_configuration=self.joint_configuration,
            joint_velocity=self.joint_velocity,
            linear_velocity_base_to_world_in_world=sampled_linear_velocity,
            orientation_base_in_world=sampled_orientation,
            position_base_in_world=sampled_position,
            randomization=self.randomization,
        )

    def to_row(self) -> np.ndarray:
        """``init_state[25]`` row of the C ABI."""
        out = np.zeros(_abi.INIT_DIM)
        out[_abi.INIT_POS:_abi.INIT_POS + 3] = self.position_base_in_world
        out[_abi.INIT_QUAT:_abi.INIT_QUAT + 4] = self.orientation_quat_wxyz
        out[_abi.INIT_LINVEL:_abi.INIT_LINVEL + 3] = self.linear_velocity_base_to_world_in_world
        out[_abi.INIT_ANGVEL:_abi.INIT_ANGVEL + 3] = self.angular_velocity_base_in_base
        out[_abi.INIT_Q:_abi.INIT_Q + 6] = self.joint_configuration
        out[_abi.INIT_QD:_abi.INIT_QD + 6] = self.joint_velocity
        return out

    def apply_to_config(self, config: _abi.UpkieSimConfig) -> None:
        """Copy the nominal pose and randomisation bounds into the simulation
        configuration (used by the on-device sampler)."""
        for k in range(3):
            config.init_position[k] = float(self.position_base_in_world[k])
            config.rand_linear_velocity[k] = float(self.randomization.linear_velocity[k])
        for k in range(4):
            config.init_quat[k] = float(self.orientation_quat_wxyz[k])
        r = self.randomization
        config.rand_roll, config.rand_pitch = float(r.roll), float(r.pitch)
        config.rand_x, config.rand_z = float(r.x), float(r.z)
        config.rand_omega_x, config.rand_omega_y = float(r.omega_x), float(r.omega_y)
        for j in range(6):
            config.init_joint_configuration[j] = float(self.joint_configuration[j])
        for k in range(3):
            config.init_angular_velocity[k] = float(self.angular_velocity_base_in_base[k])
            config.init_linear_velocity[k] = float(self.linear_velocity_base_to_world_in_world[k])

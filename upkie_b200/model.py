# SPDX-License-Identifier: Apache-2.0
"""Rigid-body model constants for the vectorised Upkie simulation.

Mirrors what ``upkie.model.Model`` provides (``upkie/model/model.py:57-110``:
``wheel_radius``, ``wheel_base``, ``left_wheeled``, ``rotation_base_to_imu``,
``joints`` in URDF order with limits) and adds the masses / inertias the
simulator needs.

**Stand-in inertias.** The reference reads every mass, inertia and joint
origin from the URDF of the un-vendored ``upkie_description`` 2.2.0 package,
which is absent from this image. ``standard_upkie()`` therefore authors a
7-body model (fixed joints lumped) that satisfies every aggregate the
reference's own tests pin:

- total mass 5.3382 kg and centre of mass (-0.0059, 0, -0.2455) in the base
  frame at the zero configuration
  (``upkie/cpp/interfaces/bullet/tests/utils_test.cpp:89-98``,
  ``upkie/cpp/interfaces/tests/BulletInterfaceTest.cpp:328-330``);
- wheel radius 0.05 m, wheel base 0.3048 m, left-wheeled,
  ``rotation_base_to_imu = diag(-1, 1, -1)`` (``tests/model/test_model.py:64-92``);
- joint limits of ``docs/kinematics.md:45-55``.

Per-link values are *parity unpinned* (SURVEY.md section 8c).
``Model.from_urdf`` overwrites them from a real ``upkie.urdf`` when one is
available.
"""

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _abi

JOINT_NAMES = _abi.JOINT_NAMES


_STANDIN_LINK_BODY = {
    "base": 0, "torso": 0, "imu": 0,
    "left_upper_leg": 1, "left_lower_leg": 2, "left_wheel_hub": 3, "left_wheel_tire": 3,
    "right_upper_leg": 4, "right_lower_leg": 5, "right_wheel_hub": 6, "right_wheel_tire": 6,
}


class ExternalForce:
    """External force applied to a robot link (``upkie/utils/external_force.py:9-43``): a force vector in
    newtons, in the world frame or (``local=True``) in the link frame."""

    def __init__(self, force, local: bool = False):
        force = np.array(force, dtype=np.float64)
        if force.shape != (3,) and not (force.ndim == 2 and force.shape[1] == 3):  # [N, 3]: one per env (vector envs)
            raise ValueError(f"Force must be a 3D vector, got shape {force.shape}")
        self.force = force
        self.local = local

    def __repr__(self) -> str:
        return f"ExternalForce(force={self.force.tolist()}, local={self.local})"


class JointProperties:
    """Per-joint simulation properties (``upkie/model/joint_properties.py:4-40``):
    kinetic friction torque and the standard deviations of the Gaussian white
    noise on the applied and on the observed torque, all in N·m."""

    def __init__(
        self,
        friction: float = 0.0,
        torque_control_noise: float = 0.0,
        torque_measurement_noise: float = 0.0,
    ):
        self.friction = friction
        self.torque_control_noise = torque_control_noise
        self.torque_measurement_noise = torque_measurement_noise


@dataclass
class JointLimit:
    """Same fields as ``upkie.model.joint_limit.JointLimit``."""

    lower: float
    upper: float
    velocity: float
    effort: float


@dataclass
class Joint:
    """Same fields as ``upkie.model.joint.Joint`` that the path uses."""

    name: str
    idx_q: int
    limit: JointLimit


@dataclass
class Model:
    """Upkie model: kinematic constants of ``upkie.model.Model`` + dynamics.

    Body ``0`` is the floating-base lump, body ``i`` (1..6) hangs from
    ``parent[i]`` through revolute joint ``i - 1``; body frames sit at their
    joint origin and are aligned with the base frame at the zero configuration.
    """

    parent: List[int]
    joint_origin: np.ndarray  # [6, 3] in the parent body frame
    joint_axis: np.ndarray  # [6, 3]
    mass: np.ndarray  # [7]
    com: np.ndarray  # [7, 3]
    inertia: np.ndarray  # [7, 6] xx, yy, zz, xy, xz, yz about the CoM
    q_lower: np.ndarray
    q_upper: np.ndarray
    qd_max: np.ndarray
    tau_max: np.ndarray
    wheel_radius: float
    wheel_base: float
    left_wheeled: bool
    imu_position: np.ndarray
    rotation_base_to_imu: np.ndarray
    rotation_ars_to_world: np.ndarray = field(
        default_factory=lambda: np.diag([1.0, -1.0, -1.0])
    )
    source: str = "stand-in"
    # URDF link name -> index of the moving body (lump) it belongs to; what PyBulletBackend keeps as
    # ``__link_index`` (pybullet_backend.py:125-147), after fixed-joint lumping
    link_body: dict = field(default_factory=lambda: dict(_STANDIN_LINK_BODY))
    # link frame -> body (lump) frame rotation at the zero configuration, identity when absent
    link_rotation: dict = field(default_factory=dict)
    # Collision shapes of the links other than the tires, reduced to what they are against the ground plane
    # (``UpkieModel.collision_*``): points in the frame of the moving body they belong to, with a radius, and the
    # name of the URDF link that carries the shape (what ``get_contact_points`` reports). Bullet collides every link
    # that has a ``<collision>`` with plane.urdf (``pybullet_backend.py:115,121``); ``upkie.model.Link`` parses the
    # same elements (``upkie/model/link.py:53-91``).
    collision_body: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=int))
    collision_point: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    collision_radius: np.ndarray = field(default_factory=lambda: np.zeros(0))
    collision_link: List[str] = field(default_factory=list)
    # the shapes themselves, for write_urdf: dicts {link, body, kind ("box" / "sphere" / "capsule" / "cylinder"), size,
    # center (body frame, base axes)}
    collision_shapes: List[dict] = field(default_factory=list)

    def add_collision_shape(self, link: str, body: int, kind: str, size, center, rotation=None) -> None:
        """Attach a collision shape to moving body ``body``: ``box`` (size = the three edge lengths) -> its 8 corners;
        ``sphere`` (size = [radius]) -> its centre; ``capsule`` / ``cylinder`` (size = [radius, length], axis along the
        shape's z) -> the two end points of the axis with the radius (for a cylinder that rounds its rims).
        ``center`` and ``rotation`` (3x3, shape axes -> body axes at the zero configuration) place the shape in the body
        frame."""
        R = np.eye(3) if rotation is None else np.asarray(rotation, dtype=float).reshape(3, 3)
        c = np.asarray(center, dtype=float)
        size = [float(x) for x in size]
        if kind == "box":
            pts = [c + R @ (0.5 * np.array([sx * size[0], sy * size[1], sz * size[2]]))
                   for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
            rad = [0.0] * 8
        elif kind == "sphere":
            pts, rad = [c], [size[0]]
        elif kind in ("capsule", "cylinder"):
            half = 0.5 * size[1] * (R @ np.array([0.0, 0.0, 1.0]))
            pts, rad = [c - half, c + half], [size[0], size[0]]
        else:
            raise ValueError(f"unsupported collision shape {kind}")
        if len(self.collision_body) + len(pts) > _abi.MAX_COLLISION_POINTS:
            raise ValueError(f"more than {_abi.MAX_COLLISION_POINTS} collision points")
        self.collision_body = np.concatenate([self.collision_body, np.full(len(pts), int(body), dtype=int)])
        self.collision_point = np.concatenate([self.collision_point.reshape(-1, 3), np.asarray(pts)])
        self.collision_radius = np.concatenate([self.collision_radius, np.asarray(rad)])
        self.collision_link = list(self.collision_link) + [link] * len(pts)
        self.collision_shapes = list(self.collision_shapes) + [
            dict(link=link, body=int(body), kind=kind, size=size, center=c.tolist(), rotation=R.tolist())]

    def external_force_rows(self, external_forces: dict, n: int = 1):
        """``{link name: ExternalForce}`` (``PyBulletBackend.set_external_forces``, ``pybullet_backend.py:603-625``)
        -> ``(force[n, 7, 3] float32, local_mask)`` for ``UpkieSim.set_external_forces``. ``force`` vectors may
        be ``[3]`` (same for every env) or ``[n, 3]``."""
        from .exceptions import UpkieRuntimeError

        rows = np.zeros((n, 7, 3), dtype=np.float32)
        frame = {}
        for name, ef in external_forces.items():
            b = self.link_body.get(name)
            if b is None:
                raise UpkieRuntimeError(f"Robot does not have a link named '{name}'")
            f = np.asarray(ef.force, dtype=np.float64)
            f = np.broadcast_to(f, (n, 3))
            local = bool(ef.local)
            if frame.setdefault(b, local) != local:
                raise UpkieRuntimeError(f"links lumped into body {b} carry forces in different frames")
            if local and name in self.link_rotation:
                f = f @ np.asarray(self.link_rotation[name], dtype=np.float64).T
            rows[:, b] += f.astype(np.float32)
        mask = sum(1 << b for b, loc in frame.items() if loc)
        return rows, mask

    # ---- upkie.model.Model API ------------------------------------------
    @property
    def joints(self) -> List[Joint]:
        """Actuated joints in URDF order (``model.py:111-114``)."""
        return [
            Joint(
                name,
                j,
                JointLimit(
                    float(self.q_lower[j]),
                    float(self.q_upper[j]),
                    float(self.qd_max[j]),
                    float(self.tau_max[j]),
                ),
            )
            for j, name in enumerate(JOINT_NAMES)
        ]

    @property
    def joint_names(self) -> set:
        return set(JOINT_NAMES)

    @property
    def upper_leg_joints(self):
        return tuple(j for j in self.joints if "hip" in j.name or "knee" in j.name)

    @property
    def wheel_joints(self):
        return tuple(j for j in self.joints if "wheel" in j.name)

    # ---- aggregates --------------------------------------------------------
    def total_mass(self) -> float:
        return float(np.sum(self.mass))

    def body_origins_zero_config(self) -> np.ndarray:
        """Body frame origins in the base frame at the zero configuration."""
        o = np.zeros((7, 3))
        for i in range(1, 7):
            o[i] = o[self.parent[i]] + self.joint_origin[i - 1]
        return o

    def com_zero_config(self) -> np.ndarray:
        o = self.body_origins_zero_config()
        return ((o + self.com) * self.mass[:, None]).sum(axis=0) / self.total_mass()

    # ---- C ABI ---------------------------------------------------------------
    def to_struct(self) -> _abi.UpkieModel:
        s = _abi.UpkieModel()
        for i in range(7):
            s.parent[i] = int(self.parent[i])
            s.mass[i] = float(self.mass[i])
            for k in range(3):
                s.com[i][k] = float(self.com[i, k])
            for k in range(6):
                s.inertia[i][k] = float(self.inertia[i, k])
        s.left_wheeled = 1 if self.left_wheeled else 0
        for j in range(6):
            for k in range(3):
                s.joint_origin[j][k] = float(self.joint_origin[j, k])
                s.joint_axis[j][k] = float(self.joint_axis[j, k])
            s.q_lower[j] = float(self.q_lower[j])
            s.q_upper[j] = float(self.q_upper[j])
            s.qd_max[j] = float(self.qd_max[j])
            s.tau_max[j] = float(self.tau_max[j])
        s.wheel_radius = float(self.wheel_radius)
        s.wheel_base = float(self.wheel_base)
        for k in range(3):
            s.imu_position[k] = float(self.imu_position[k])
        R = np.asarray(self.rotation_base_to_imu, dtype=float).reshape(9)
        for k in range(9):
            s.rotation_base_to_imu[k] = float(R[k])
        s.n_collision_points = len(self.collision_body)
        for p in range(len(self.collision_body)):
            s.collision_body[p] = int(self.collision_body[p])
            s.collision_radius[p] = float(self.collision_radius[p])
            for k in range(3):
                s.collision_point[p][k] = float(self.collision_point[p][k])
        return s

    # ---- constructors ---------------------------------------------------------
    @staticmethod
    def standard_upkie() -> "Model":
        """Stand-in for ``upkie_description`` (see module docstring)."""
        half_base = 0.3048 / 2.0
        y_hip, y_step = 0.1, (half_base - 0.1) / 2.0
        z_hip, l_upper, l_lower = -0.19, 0.17, 0.17
        joint_origin = np.array(
            [
                [0.0, +y_hip, z_hip],
                [0.0, +y_step, -l_upper],
                [0.0, +y_step, -l_lower],
                [0.0, -y_hip, z_hip],
                [0.0, -y_step, -l_upper],
                [0.0, -y_step, -l_lower],
            ]
        )
        # moteus convention: the left-wheel hub z-axis points along -y of the
        # base and the joint axis is -z of the joint frame (docs/kinematics.md
        # "Link and joint names"), i.e. +y in base coordinates: positive
        # left-wheel velocity rolls forward (left_wheeled, model.py:104).
        joint_axis = np.array(
            [[0, 1, 0], [0, 1, 0], [0, 1, 0], [0, -1, 0], [0, -1, 0], [0, -1, 0]],
            dtype=float,
        )
        m_upper, m_lower, m_wheel = 0.70, 0.35, 0.30
        total_mass = 5.3382
        target_com = np.array([-0.0059, 0.0, -0.2455])
        mass = np.array(
            [0.0, m_upper, m_lower, m_wheel, m_upper, m_lower, m_wheel]
        )
        mass[0] = total_mass - mass.sum()
        com = np.array(
            [
                [0.0, 0.0, 0.0],
                [0.0, +0.02, -0.10],
                [0.0, +0.015, -0.10],
                [0.0, 0.0, 0.0],
                [0.0, -0.02, -0.10],
                [0.0, -0.015, -0.10],
                [0.0, 0.0, 0.0],
            ]
        )
        parent = [-1, 0, 1, 2, 0, 4, 5]
        # solve the base-lump CoM so that the aggregate CoM is the pinned one
        o = np.zeros((7, 3))
        for i in range(1, 7):
            o[i] = o[parent[i]] + joint_origin[i - 1]
        legs = ((o[1:] + com[1:]) * mass[1:, None]).sum(axis=0)
        com[0] = (total_mass * target_com - legs) / mass[0]
        bx, by, bz = 0.14, 0.22, 0.22  # torso box
        inertia = np.array(
            [
                [
                    mass[0] / 12.0 * (by**2 + bz**2),
                    mass[0] / 12.0 * (bx**2 + bz**2),
                    mass[0] / 12.0 * (bx**2 + by**2),
                    0.0,
                    0.0,
                    0.0,
                ],
                [3.0e-3, 3.0e-3, 8.0e-4, 0, 0, 0],
                [1.4e-3, 1.4e-3, 3.0e-4, 0, 0, 0],
                [3.2e-4, 6.0e-4, 3.2e-4, 0, 0, 0],
                [3.0e-3, 3.0e-3, 8.0e-4, 0, 0, 0],
                [1.4e-3, 1.4e-3, 3.0e-4, 0, 0, 0],
                [3.2e-4, 6.0e-4, 3.2e-4, 0, 0, 0],
            ]
        )
        inf = math.inf
        model = Model(
            parent=parent,
            joint_origin=joint_origin,
            joint_axis=joint_axis,
            mass=mass,
            com=com,
            inertia=inertia,
            q_lower=np.array([-1.26, -2.51, -inf, -1.26, -2.51, -inf]),
            q_upper=np.array([+1.26, +2.51, +inf, +1.26, +2.51, +inf]),
            qd_max=np.array([28.8, 28.8, 111.0, 28.8, 28.8, 111.0]),
            tau_max=np.array([16.0, 16.0, 1.7, 16.0, 16.0, 1.7]),
            wheel_radius=0.05,
            wheel_base=2.0 * half_base,
            left_wheeled=True,
            imu_position=np.array([-0.01, 0.0, -0.06]),
            rotation_base_to_imu=np.diag([-1.0, 1.0, -1.0]),
        )
        # collision box of the torso: the box the base lump's inertia was authored from, about its centre of mass
        # (stand-in like the inertias: the real shapes live in upkie_description)
        model.add_collision_shape("torso", 0, "box", (bx, by, bz), com[0])
        return model

    @staticmethod
    def from_urdf(urdf_path: str) -> "Model":
        """Build the 7-body model from a real ``upkie.urdf``.

        Follows the parsing conventions of ``upkie/model/kinematic_tree.py:52-143``
        (URDF roll-pitch-yaw -> rotation, joints in file order) and adds an
        ``<inertial>`` parser; links attached by fixed joints are lumped into
        their moving ancestor, frames are re-expressed aligned with the base at
        the zero configuration.
        """
        from .urdf import load_urdf_model  # local import: optional path

        return load_urdf_model(urdf_path)


def default_model(urdf_path: Optional[str] = None) -> Model:
    """``Model(urdf_path=None)`` of the reference defaults to
    ``upkie_description.URDF_PATH`` (``model.py:63-65``); here it is the real
    URDF when ``upkie_description`` is importable, else the stand-in."""
    if urdf_path is not None:
        return Model.from_urdf(urdf_path)
    try:
        import upkie_description  # type: ignore

        return Model.from_urdf(upkie_description.URDF_PATH)
    except ImportError:
        return Model.standard_upkie()

@dataclass(eq=False)
class PointContact:
    """Same constructor, attributes and ``repr`` as ``upkie.utils.point_contact.PointContact``
    (``upkie/utils/point_contact.py:8-55``): one contact seen from the robot's side."""

    link_name: str
    position_contact_in_world: np.ndarray
    force_in_world: np.ndarray

    def __repr__(self) -> str:
        fields = (
            f"link_name='{self.link_name}'",
            f"position_contact_in_world={np.asarray(self.position_contact_in_world).tolist()}",
            f"force_in_world={np.asarray(self.force_in_world).tolist()}",
        )
        return "PointContact(" + ", ".join(fields) + ")"


def wheel_contact_points(model, state_row, substep_dt: float, breaking_threshold: float = 0.02):
    """Contact points of the two tires for one robot, from a simulator state row (``UPKIE_ST_*`` layout).

    Restates the collision pass of the step kernel on the host (``physics_substep_paired``, sim_pair.cuh: the
    lowest point of the tire circle against the plane z = 0, a contact while it is closer than Bullet's contact
    breaking threshold) and reads the normal impulses the last substep applied (``UPKIE_ST_CONTACT_IMPULSE``).
    Returns ``[(side, position_in_world[3], force_in_world[3]), ...]``; ``side`` is 0 (left) or 1 (right). The force is
    the one the ground exerted on the tire during the last substep: normal impulse along +z and the two friction
    impulses (``UPKIE_ST_FRICTION_IMPULSE``) along the contact's rolling and lateral directions - what
    ``getContactPoints`` reports as normalForce, lateralFriction1 / 2 (``pybullet_backend.py:696-709``)."""
    row = np.asarray(state_row, dtype=float)
    pos = row[_abi.ST_POS:_abi.ST_POS + 3]
    w, x, y, z = row[_abi.ST_QUAT:_abi.ST_QUAT + 4]
    R = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])
    q = row[_abi.ST_Q:_abi.ST_Q + 6]
    zb = R.T @ np.array([0.0, 0.0, 1.0])  # world up in base coordinates
    out = []
    for side in (0, 1):
        origin, phi = np.zeros(3), 0.0
        for k in range(3):  # hip, knee, wheel: every joint turns about +-y of the base
            j = 3 * side + k
            c, s = np.cos(phi), np.sin(phi)
            origin = origin + np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]]) @ np.asarray(model.joint_origin[j], dtype=float)
            phi += float(model.joint_axis[j][1]) * q[j]
        n_xz = np.hypot(zb[0], zb[2])
        if n_xz <= 1e-6:  # wheel lying flat: no rim contact
            continue
        down = np.array([-zb[0], 0.0, -zb[2]]) / n_xz  # in the wheel plane, towards the ground
        p_base = origin + float(model.wheel_radius) * down
        p_world = pos + R @ p_base
        if p_world[2] >= breaking_threshold:
            continue
        impulse = float(row[_abi.ST_CONTACT_IMPULSE + side])
        # friction directions of the kernels (sim_pair.cuh): t1 = s (zb.z, 0, -zb.x) / |.| (rolling), t2 = s (zb x t1)
        # (lateral), s = sign of the wheel's joint axis; both lie in the ground plane
        sgn = float(model.joint_axis[3 * side + 2][1])
        t1 = sgn * np.array([zb[2], 0.0, -zb[0]]) / n_xz
        t2 = np.cross(zb, t1)
        lam_t = row[_abi.ST_FRICTION_IMPULSE + 2 * side:_abi.ST_FRICTION_IMPULSE + 2 * side + 2] if len(row) > _abi.ST_FRICTION_IMPULSE else (0.0, 0.0)
        force = (impulse * np.array([0.0, 0.0, 1.0]) + R @ (float(lam_t[0]) * t1 + float(lam_t[1]) * t2)) / substep_dt
        out.append((side, p_world, force))
    return out

TIRE_LINKS = ("left_wheel_tire", "right_wheel_tire")


def body_points_in_world(model, state_row):
    """World positions ``[K, 3]`` of the model's collision points (``Model.collision_point``) for one robot, from its
    state row: the kinematics of the step kernels restated on the host (every joint turns about +-y of the base)."""
    row = np.asarray(state_row, dtype=float)
    pos = row[_abi.ST_POS:_abi.ST_POS + 3]
    w, x, y, z = row[_abi.ST_QUAT:_abi.ST_QUAT + 4]
    R = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])
    q = row[_abi.ST_Q:_abi.ST_Q + 6]
    frames = [(np.zeros(3), np.eye(3))]  # body -> (origin, rotation) in base coordinates
    for side in (0, 1):
        origin, phi = np.zeros(3), 0.0
        for k in range(3):
            j = 3 * side + k
            c, s = np.cos(phi), np.sin(phi)
            Ry = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
            origin = origin + Ry @ np.asarray(model.joint_origin[j], dtype=float)
            phi += float(model.joint_axis[j][1]) * q[j]
            c, s = np.cos(phi), np.sin(phi)
            frames.append((origin, np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])))
    out = np.zeros((len(model.collision_body), 3))
    for p, b in enumerate(model.collision_body):
        o, Rb = frames[int(b)]
        out[p] = pos + R @ (o + Rb @ np.asarray(model.collision_point[p], dtype=float))
    return out


def body_contact_points(model, state_row, body_rec_row, substep_dt: float):
    """Body-ground contacts of one robot from the record of the last substep (``UpkieSim.get_body_contacts`` row,
    layout ``upkie_b200_get_body_contacts``): ``[(link name, position_in_world[3], force_in_world[3]), ...]``. The
    position is the lowest point of the collision sphere at the current state; the force sums the normal impulse
    (+z) and the two friction impulses (world -y and +x, the solver's directions), over the substep."""
    rec = np.asarray(body_rec_row, dtype=float)
    mask = int(round(rec[0]))
    if mask == 0 or len(model.collision_body) == 0:
        return []
    pts = body_points_in_world(model, state_row)
    out = []
    for c in range(min(bin(mask).count("1"), _abi.MAX_BODY_CONTACTS)):  # one slot per bit of the mask, in point order
        p = int(round(rec[1 + 4 * c]))
        lam_n, lam_1, lam_2 = rec[1 + 4 * c + 1:1 + 4 * c + 4]
        position = pts[p] - np.array([0.0, 0.0, float(model.collision_radius[p])])
        force = np.array([lam_2, -lam_1, lam_n]) / substep_dt
        out.append((model.collision_link[p], position, force))
    return out


def contact_points_from_state(model, state_row, config, link_name=None, body_rec_row=None):
    """``PyBulletBackend.get_contact_points`` (``pybullet_backend.py:660-716``) for one robot from its state row:
    ``PointContact`` instances on ``left_wheel_tire`` / ``right_wheel_tire`` and, with ``body_rec_row`` (the robot's row
    of ``UpkieSim.get_body_contacts``), on the links that carry collision shapes (the torso...), filtered by
    ``link_name``. A link without simulated contacts, or one the robot does not have, yields ``[]`` like the reference.
    ``force_in_world`` = normal + two friction components of the last substep, as the reference sums them
    (``pybullet_backend.py:696-709``)."""
    h = float(config.dt) / int(config.nb_substeps)
    out = []
    if link_name is None or link_name in TIRE_LINKS:
        contacts = wheel_contact_points(model, state_row, h, breaking_threshold=float(config.contact_breaking_threshold))
        out += [
            PointContact(TIRE_LINKS[side], position, np.asarray(force, dtype=float))
            for side, position, force in contacts
            if link_name is None or TIRE_LINKS[side] == link_name
        ]
    if body_rec_row is not None and (link_name is None or link_name not in TIRE_LINKS):
        out += [
            PointContact(name, position, force)
            for name, position, force in body_contact_points(model, state_row, body_rec_row, h)
            if link_name is None or name == link_name
        ]
    return out

# SPDX-License-Identifier: Apache-2.0
"""URDF -> 7-body ``Model`` (fixed joints lumped, frames aligned with the base).

Follows the parsing conventions of the reference's kinematic tree
(``upkie/model/kinematic_tree.py:52-143``: links and joints in file order,
URDF roll-pitch-yaw -> rotation as ``upkie/utils/rotations.py:74-102``, joint
axis default ``(1, 0, 0)``, limits from ``<limit>``) and of ``Model``
(``upkie/model/model.py:57-144``: wheel radius from the ``left_wheel_tire``
collision cylinder, wheel base from the tire frames, left-wheeledness from the
``left_wheel_hub`` z-axis, ``rotation_base_to_imu``), and adds what the
simulator needs: ``<inertial>`` blocks, lumped per moving body.
"""

import math
import xml.etree.ElementTree as ET
from typing import Dict, List, Tuple

import numpy as np

from . import _abi
from .exceptions import ModelError


def rotation_matrix_from_rpy(rpy) -> np.ndarray:
    """``R = Rz(yaw) Ry(pitch) Rx(roll)`` (``upkie/utils/rotations.py:74-102``)."""
    roll, pitch, yaw = rpy
    cr, sr = math.cos(roll), math.sin(roll)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cy, sy = math.cos(yaw), math.sin(yaw)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ]
    )


def _vec(s, default):
    return np.array([float(x) for x in s.split()]) if s else np.array(default, dtype=float)


def _origin(elem) -> Tuple[np.ndarray, np.ndarray]:
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return np.eye(3), np.zeros(3)
    return rotation_matrix_from_rpy(_vec(o.get("rpy"), [0, 0, 0])), _vec(o.get("xyz"), [0, 0, 0])


def load_urdf_model(urdf_path: str):
    from .model import Model

    root = ET.parse(urdf_path).getroot()
    links: Dict[str, ET.Element] = {l.get("name"): l for l in root.findall("link")}
    joints: List[ET.Element] = root.findall("joint")
    child_of: Dict[str, List[ET.Element]] = {}
    has_parent = set()
    for j in joints:
        child_of.setdefault(j.find("parent").get("link"), []).append(j)
        has_parent.add(j.find("child").get("link"))
    roots = [n for n in links if n not in has_parent]
    if len(roots) != 1:
        raise ModelError(f"URDF should have exactly one root link, found {roots}")
    base = roots[0]

    # zero-configuration transform of every link frame to the base frame
    T: Dict[str, Tuple[np.ndarray, np.ndarray]] = {base: (np.eye(3), np.zeros(3))}
    joint_frame: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}
    body_of: Dict[str, int] = {base: 0}  # link -> moving body index
    joint_of_body: Dict[int, ET.Element] = {}
    parent_body: Dict[int, int] = {0: -1}
    stack = [base]
    while stack:
        parent = stack.pop()
        Rp, pp = T[parent]
        for j in child_of.get(parent, []):
            Rj, pj = _origin(j)
            child = j.find("child").get("link")
            T[child] = (Rp @ Rj, pp + Rp @ pj)
            joint_frame[j.get("name")] = T[child]
            if j.get("type") in ("revolute", "continuous") and j.get("name") in _abi.JOINT_NAMES:
                b = _abi.JOINT_NAMES.index(j.get("name")) + 1
                body_of[child] = b
                joint_of_body[b] = j
                parent_body[b] = body_of[parent]
            elif j.get("type") == "fixed":
                body_of[child] = body_of[parent]
            else:
                raise ModelError(f"unsupported joint {j.get('name')} of type {j.get('type')}")
            stack.append(child)
    if sorted(joint_of_body) != [1, 2, 3, 4, 5, 6]:
        raise ModelError(f"URDF must have the six joints {_abi.JOINT_NAMES}")
    parent = [parent_body[b] for b in range(7)]
    if parent != [-1, 0, 1, 2, 0, 4, 5]:
        raise ModelError(f"unexpected kinematic tree {parent}")

    body_origin = np.zeros((7, 3))
    joint_axis = np.zeros((6, 3))
    q_lower, q_upper = np.full(6, -math.inf), np.full(6, math.inf)
    qd_max, tau_max = np.zeros(6), np.zeros(6)
    for b in range(1, 7):
        j = joint_of_body[b]
        Rj, pj = joint_frame[j.get("name")]
        body_origin[b] = pj
        axis = j.find("axis")
        a = _vec(axis.get("xyz") if axis is not None else None, [1, 0, 0])
        joint_axis[b - 1] = Rj @ (a / np.linalg.norm(a))
        lim = j.find("limit")
        if lim is not None:
            if j.get("type") == "revolute":
                q_lower[b - 1] = float(lim.get("lower", -math.inf))
                q_upper[b - 1] = float(lim.get("upper", math.inf))
            qd_max[b - 1] = float(lim.get("velocity", 0.0))
            tau_max[b - 1] = float(lim.get("effort", 0.0))
    joint_origin = np.array([body_origin[b] - body_origin[parent[b]] for b in range(1, 7)])

    # lump inertials per moving body (all expressed in base axes)
    mass = np.zeros(7)
    first = np.zeros((7, 3))
    parts: List[Tuple[int, float, np.ndarray, np.ndarray]] = []
    for name, link in links.items():
        inertial = link.find("inertial")
        if inertial is None or name not in body_of:
            continue
        m = float(inertial.find("mass").get("value"))
        Ri, pi = _origin(inertial)
        Rl, pl = T[name]
        c = pl + Rl @ pi
        it = inertial.find("inertia")
        I = np.zeros((3, 3))
        if it is not None:
            ixx, iyy, izz = float(it.get("ixx", 0)), float(it.get("iyy", 0)), float(it.get("izz", 0))
            ixy, ixz, iyz = float(it.get("ixy", 0)), float(it.get("ixz", 0)), float(it.get("iyz", 0))
            I = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
        R = Rl @ Ri
        b = body_of[name]
        parts.append((b, m, c, R @ I @ R.T))
        mass[b] += m
        first[b] += m * c
    if np.any(mass <= 0):
        raise ModelError("every moving body needs a positive mass")
    com_base = first / mass[:, None]
    inertia = np.zeros((7, 3, 3))
    for b, m, c, I in parts:
        d = c - com_base[b]
        inertia[b] += I + m * (d @ d * np.eye(3) - np.outer(d, d))
    com = com_base - body_origin
    inertia6 = np.array([[I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]] for I in inertia])

    # Model attributes (upkie/model/model.py:66-110)
    for needed in ("left_wheel_tire", "right_wheel_tire", "left_wheel_hub", "imu"):
        if needed not in T:
            raise ModelError(f"{needed} link not found in URDF")
    pos_left, pos_right = T["left_wheel_tire"][1], T["right_wheel_tire"][1]
    wheel_base = float(np.linalg.norm(pos_left - pos_right))
    geoms = links["left_wheel_tire"].findall("collision")
    if len(geoms) != 1 or geoms[0].find("geometry").find("cylinder") is None:
        raise ModelError("left_wheel_tire should have exactly one collision geometry, a cylinder")
    wheel_radius = float(geoms[0].find("geometry").find("cylinder").get("radius"))
    z_hub = T["left_wheel_hub"][0][:, 2]
    if abs(z_hub[0]) > 1e-4 or abs(z_hub[2]) > 1e-4:
        raise ModelError(f"the z-axis of the left-wheel hub {z_hub} is not aligned with the y-axis of the base frame")
    R_base_from_imu, p_imu = T["imu"]
    # collision shapes of the links other than the tires -> collision points of the moving bodies
    shapes = []
    for name, link in links.items():
        if name not in body_of or name.endswith("_wheel_tire"):
            continue
        for col in link.findall("collision"):
            geom = col.find("geometry")
            if geom is None:
                continue
            Rc, pc = _origin(col)
            Rl, pl = T[name]
            b = body_of[name]
            center = pl + Rl @ pc - body_origin[b]
            rot = Rl @ Rc
            if geom.find("box") is not None:
                shapes.append((name, b, "box", _vec(geom.find("box").get("size"), [0, 0, 0]), center, rot))
            elif geom.find("sphere") is not None:
                shapes.append((name, b, "sphere", [float(geom.find("sphere").get("radius"))], center, rot))
            elif geom.find("cylinder") is not None or geom.find("capsule") is not None:
                kind = "cylinder" if geom.find("cylinder") is not None else "capsule"
                g = geom.find(kind)
                shapes.append((name, b, kind, [float(g.get("radius")), float(g.get("length"))], center, rot))
            # meshes: no analytic reduction to points; such a link does not collide here
    model = Model(
        parent=parent,
        joint_origin=joint_origin,
        joint_axis=joint_axis,
        mass=mass,
        com=com,
        inertia=inertia6,
        q_lower=q_lower,
        q_upper=q_upper,
        qd_max=qd_max,
        tau_max=tau_max,
        wheel_radius=wheel_radius,
        wheel_base=wheel_base,
        left_wheeled=bool(z_hub[1] < 0),
        imu_position=p_imu.copy(),
        rotation_base_to_imu=R_base_from_imu.T.copy(),
        source=urdf_path,
        link_body=dict(body_of),
        link_rotation={name: T[name][0].copy() for name in body_of},
    )
    for name, b, kind, size, center, rot in shapes:
        try:
            model.add_collision_shape(name, b, kind, size, center, rot)
        except ValueError as exc:  # more points than UpkieModel carries: the remaining shapes do not collide
            import warnings

            warnings.warn(f"collision shape of link '{name}' ignored ({exc})")
    return model


def write_urdf(model, path: str, split_fixed_links: bool = True) -> None:
    """Write ``model`` as a URDF (used by the tests to round-trip the loader and as a
    template for users without ``upkie_description``). With ``split_fixed_links`` the
    base lump is split into ``base`` + ``torso`` + ``imu`` and each wheel into hub +
    tire, attached by fixed joints, as in the real description."""
    def inertial(m, c, I6):
        return (
            f'<inertial><origin xyz="{c[0]:.17g} {c[1]:.17g} {c[2]:.17g}" rpy="0 0 0"/><mass value="{m:.17g}"/>'
            f'<inertia ixx="{I6[0]:.17g}" iyy="{I6[1]:.17g}" izz="{I6[2]:.17g}" ixy="{I6[3]:.17g}" ixz="{I6[4]:.17g}" iyz="{I6[5]:.17g}"/></inertial>'
        )

    def rpy_of(Rm):
        return (math.atan2(Rm[2, 1], Rm[2, 2]), math.atan2(-Rm[2, 0], math.hypot(Rm[0, 0], Rm[1, 0])),
                math.atan2(Rm[1, 0], Rm[0, 0]))

    def collisions(body, link_origin, link_rotation):
        """<collision> elements of the shapes on moving body ``body``, in a link frame at ``link_origin`` (body frame)
        with axes ``link_rotation`` (link -> body)."""
        el = []
        for sh in getattr(model, "collision_shapes", []):
            if sh["body"] != body:
                continue
            c = link_rotation.T @ (np.asarray(sh["center"], dtype=float) - link_origin)
            r, p, y = rpy_of(link_rotation.T @ np.asarray(sh["rotation"], dtype=float))
            size = sh["size"]
            if sh["kind"] == "box":
                g = f'<box size="{size[0]:.17g} {size[1]:.17g} {size[2]:.17g}"/>'
            elif sh["kind"] == "sphere":
                g = f'<sphere radius="{size[0]:.17g}"/>'
            else:
                g = f'<{sh["kind"]} radius="{size[0]:.17g}" length="{size[1]:.17g}"/>'
            el.append(f'<collision><origin xyz="{c[0]:.17g} {c[1]:.17g} {c[2]:.17g}" rpy="{r:.17g} {p:.17g} {y:.17g}"/>'
                      f'<geometry>{g}</geometry></collision>')
        return "".join(el)

    out = ['<?xml version="1.0"?>', '<robot name="upkie">']
    m0, c0, I0 = float(model.mass[0]), model.com[0], model.inertia[0]
    if split_fixed_links:
        # torso carries everything but two virtual grams (docs/kinematics.md "Virtual links")
        m_virtual = 0.001
        m_t = m0 - 2 * m_virtual
        # base virtual mass at the base origin, imu virtual mass at the imu frame: solve the torso CoM/inertia
        p_imu = np.asarray(model.imu_position, dtype=float)
        c_t = (m0 * c0 - m_virtual * np.zeros(3) - m_virtual * p_imu) / m_t
        I = np.array([[I0[0], I0[3], I0[4]], [I0[3], I0[1], I0[5]], [I0[4], I0[5], I0[2]]])

        def shift(m, d):
            return m * (d @ d * np.eye(3) - np.outer(d, d))

        I_t = I - shift(m_virtual, -c0) - shift(m_virtual, p_imu - c0) - shift(m_t, c_t - c0)
        torso_origin = np.array([0.0, 0.0, -0.1])  # tests/model/test_kinematic_tree.py:31-36
        out.append(f'<link name="base">{inertial(m_virtual, np.zeros(3), np.zeros(6))}</link>')
        It6 = [I_t[0, 0], I_t[1, 1], I_t[2, 2], I_t[0, 1], I_t[0, 2], I_t[1, 2]]
        out.append(f'<link name="torso">{inertial(m_t, c_t - torso_origin, It6)}{collisions(0, torso_origin, np.eye(3))}</link>')
        out.append(f'<joint name="torso_fix" type="fixed"><parent link="base"/><child link="torso"/>'
                   f'<origin xyz="0 0 -0.1" rpy="0 0 0"/></joint>')
        # imu frame: rotation_base_to_imu^T = R_base_from_imu; diag(-1, 1, -1) = rotation of pi about y
        R = np.asarray(model.rotation_base_to_imu, dtype=float).reshape(3, 3).T
        pitch = math.atan2(-R[2, 0], math.hypot(R[0, 0], R[1, 0]))
        yaw = math.atan2(R[1, 0], R[0, 0])
        roll = math.atan2(R[2, 1], R[2, 2])
        pi_t = p_imu - torso_origin
        out.append(f'<link name="imu">{inertial(m_virtual, np.zeros(3), np.zeros(6))}</link>')
        out.append(f'<joint name="imu_fix" type="fixed"><parent link="torso"/><child link="imu"/>'
                   f'<origin xyz="{pi_t[0]:.17g} {pi_t[1]:.17g} {pi_t[2]:.17g}" rpy="{roll:.17g} {pitch:.17g} {yaw:.17g}"/></joint>')
    else:
        out.append(f'<link name="base">{inertial(m0, c0, I0)}{collisions(0, np.zeros(3), np.eye(3))}</link>')
        R = np.asarray(model.rotation_base_to_imu, dtype=float).reshape(3, 3).T
        pitch = math.atan2(-R[2, 0], math.hypot(R[0, 0], R[1, 0]))
        yaw = math.atan2(R[1, 0], R[0, 0])
        roll = math.atan2(R[2, 1], R[2, 2])
        p_imu = np.asarray(model.imu_position, dtype=float)
        out.append('<link name="imu"/>')
        out.append(f'<joint name="imu_fix" type="fixed"><parent link="base"/><child link="imu"/>'
                   f'<origin xyz="{p_imu[0]:.17g} {p_imu[1]:.17g} {p_imu[2]:.17g}" rpy="{roll:.17g} {pitch:.17g} {yaw:.17g}"/></joint>')
    names = {1: "left_upper_leg", 2: "left_lower_leg", 3: "left_wheel_hub", 4: "right_upper_leg", 5: "right_lower_leg",
             6: "right_wheel_hub"}
    parent_name = {0: "torso" if split_fixed_links else "base"}
    parent_name.update(names)
    for b in range(1, 7):
        j = b - 1
        wheel = b in (3, 6)
        # moteus convention: the joint axis is -z of the joint frame (docs/kinematics.md); the joint frame is
        # rotated so that its -z is the model's axis: axis +y -> frame z = -y -> roll = +pi/2
        ay = float(model.joint_axis[j][1])
        roll = math.pi / 2 if ay > 0 else -math.pi / 2
        Rj = rotation_matrix_from_rpy((roll, 0.0, 0.0))
        po = np.asarray(model.joint_origin[j], dtype=float)
        if parent_name[model.parent[b]] == "torso" and model.parent[b] == 0:
            po = po - np.array([0.0, 0.0, -0.1])
        elif model.parent[b] != 0:
            # parent link frame is rotated like its own joint frame
            ayp = float(model.joint_axis[model.parent[b] - 1][1])
            Rp = rotation_matrix_from_rpy((math.pi / 2 if ayp > 0 else -math.pi / 2, 0.0, 0.0))
            po = Rp.T @ po
            Rj = Rp.T @ Rj
        rr = math.atan2(Rj[2, 1], Rj[2, 2])
        c_link = Rj_full(model, b).T @ np.asarray(model.com[b], dtype=float)
        Ib = model.inertia[b]
        Im = np.array([[Ib[0], Ib[3], Ib[4]], [Ib[3], Ib[1], Ib[5]], [Ib[4], Ib[5], Ib[2]]])
        Il = Rj_full(model, b).T @ Im @ Rj_full(model, b)
        I6 = [Il[0, 0], Il[1, 1], Il[2, 2], Il[0, 1], Il[0, 2], Il[1, 2]]
        out.append(f'<link name="{names[b]}">{inertial(float(model.mass[b]), c_link, I6)}'
                   f'{collisions(b, np.zeros(3), Rj_full(model, b))}</link>')
        jtype = "continuous" if wheel else "revolute"
        lim = (f'<limit effort="{float(model.tau_max[j]):.17g}" velocity="{float(model.qd_max[j]):.17g}"'
               + ("" if wheel else f' lower="{float(model.q_lower[j]):.17g}" upper="{float(model.q_upper[j]):.17g}"') + "/>")
        out.append(f'<joint name="{_abi.JOINT_NAMES[j]}" type="{jtype}"><parent link="{parent_name[model.parent[b]]}"/>'
                   f'<child link="{names[b]}"/><origin xyz="{po[0]:.17g} {po[1]:.17g} {po[2]:.17g}" rpy="{rr:.17g} 0 0"/>'
                   f'<axis xyz="0 0 -1"/>{lim}</joint>')
        if wheel:
            side = "left" if b == 3 else "right"
            out.append(f'<link name="{side}_wheel_tire"><collision><geometry><cylinder radius="{float(model.wheel_radius):.17g}" '
                       f'length="0.02"/></geometry></collision></link>')
            out.append(f'<joint name="{side}_wheel_tire_fix" type="fixed"><parent link="{names[b]}"/>'
                       f'<child link="{side}_wheel_tire"/></joint>')
            if split_fixed_links:
                # massless frames of the real description (tests/model/test_model.py:32-45): the wheel centre and
                # the ground contact point below it at the zero configuration
                down = Rj_full(model, b).T @ np.array([0.0, 0.0, -float(model.wheel_radius)])
                out.append(f'<link name="{side}_wheel_center"/>')
                out.append(f'<joint name="{side}_wheel_center_fix" type="fixed"><parent link="{side}_wheel_tire"/>'
                           f'<child link="{side}_wheel_center"/></joint>')
                out.append(f'<link name="{side}_contact"/>')
                out.append(f'<joint name="{side}_contact_fix" type="fixed"><parent link="{side}_wheel_tire"/>'
                           f'<child link="{side}_contact"/><origin xyz="{down[0]:.17g} {down[1]:.17g} {down[2]:.17g}" '
                           f'rpy="0 0 0"/></joint>')
        if split_fixed_links and b in (1, 4):
            # the hip actuator's stator: a massless frame fixed to the torso at the hip joint
            side = "left" if b == 1 else "right"
            out.append(f'<link name="{side}_hip_qdd100_stator"/>')
            out.append(f'<joint name="{side}_hip_qdd100_stator_fix" type="fixed"><parent link="torso"/>'
                       f'<child link="{side}_hip_qdd100_stator"/><origin xyz="{po[0]:.17g} {po[1]:.17g} {po[2]:.17g}" '
                       f'rpy="0 0 0"/></joint>')
    out.append("</robot>")
    with open(path, "w") as f:
        f.write("\n".join(out))


def Rj_full(model, b) -> np.ndarray:
    """Rotation base <- link frame of moving body ``b`` in the URDF written by ``write_urdf``."""
    ay = float(model.joint_axis[b - 1][1])
    return rotation_matrix_from_rpy((math.pi / 2 if ay > 0 else -math.pi / 2, 0.0, 0.0))

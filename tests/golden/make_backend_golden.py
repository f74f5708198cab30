#!/usr/bin/env python
# SPDX-License-Identifier: Apache-2.0
"""Golden runs of the reference's OWN PyBulletBackend class with a stand-in ``pybullet`` whose physics is the oracle.

Run in the build container:  python tests/golden/make_backend_golden.py

``upkie/envs/backends/pybullet_backend.py`` is imported unmodified. ``pybullet`` itself is absent, so a stand-in module
implements the ~20 API functions the backend calls (TEST INFRASTRUCTURE) on top of ONE robot of oracle/: the rigid-body
state lives in the oracle, ``stepSimulation()`` is ``oracle_substep(tau, h)``, the getters are plain kinematics on that
state. What runs as the reference's code, line for line: the substep loop and the moteus torque law with friction
(pybullet_backend.py:269-311,492-553), the reset sequence with its quirks (body-frame angular velocity handed over as
world-frame, one un-actuated substep, IMU finite-difference state and last torques not cleared, :220-267), the
observation assembly (base orientation, IMU in the ARS frame with finite-difference acceleration, floor contact,
servo block, wheel odometry, :313-490), ``randomize_inertias`` (:555-601) and ``set_external_forces`` (:603-658).
The golden spine observations pin the oracle's restatement of those rows (a4, a5, a7, a8, a9 of SURVEY.md section 8);
the physics inside ``stepSimulation`` (a6, Bullet itself) is the oracle on both sides and stays "parity unpinned".

Output: tests/golden/backend_runs.json, replayed by tests/test_backend_golden.py.
"""
import json
import os
import re
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "backend_runs.json")
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def make_fake_pybullet(b200_model, urdf_path, joint_limits=None):
    """``pybullet`` API subset over one oracle robot. Link / joint indices follow the URDF joint order like Bullet."""
    from oracle import oracle as O
    from upkie_b200 import _abi

    text = open(urdf_path).read()
    joints = re.findall(r'<joint name="([^"]+)" type="([^"]+)"><parent link="([^"]+)"/><child link="([^"]+)"', text)
    assert joints, "unexpected URDF formatting"
    pb = types.ModuleType("pybullet")
    pb.GUI, pb.DIRECT = 1, 2
    pb.COV_ENABLE_GUI, pb.COV_ENABLE_RENDERING, pb.COV_ENABLE_SHADOWS = 1, 2, 3
    pb.VELOCITY_CONTROL, pb.TORQUE_CONTROL = 0, 2
    pb.LINK_FRAME, pb.WORLD_FRAME = 1, 2
    S = types.SimpleNamespace(sim=None, h=None, gravity=None, tau=np.zeros(6), ext={}, eps=np.zeros(6), calls=[])
    joint_of_name = {name: k for k, name in enumerate(_abi.JOINT_NAMES)}
    body_of_link = dict(b200_model.link_body)

    def ensure():
        if S.sim is None:
            cfg = _abi.default_sim_config()
            cfg.gravity = -S.gravity[2]
            if joint_limits is not None:
                cfg.joint_limits = int(joint_limits)
            S.cfg = cfg
            S.sim = O.OracleSim(b200_model, cfg, 1, threads=1)

    def state():
        return S.sim.get_state()[0]

    def put(st):
        S.sim.set_state(st.reshape(1, -1))

    def R_of(quat_wxyz):
        w, x, y, z = quat_wxyz
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    pb.connect = lambda mode: 0
    pb.disconnect = lambda *a, **k: None
    pb.configureDebugVisualizer = lambda *a, **k: None
    pb.resetDebugVisualizerCamera = lambda *a, **k: None
    pb.setAdditionalSearchPath = lambda p: None
    pb.setRealTimeSimulation = lambda f: None

    def setGravity(x, y, z):
        S.gravity = (x, y, z)

    def setTimeStep(h):
        S.h = h

    def loadURDF(path, basePosition=None, baseOrientation=None):
        if os.path.basename(path) == "plane.urdf":
            return 0
        ensure()
        return 1

    pb.setGravity, pb.setTimeStep, pb.loadURDF = setGravity, setTimeStep, loadURDF
    pb.getNumJoints = lambda robot: len(joints)

    def getJointInfo(robot, idx):
        name, jtype, parent, child = joints[idx]
        info = [idx, name.encode(), {"revolute": 0, "continuous": 0, "fixed": 4}[jtype]] + [0] * 9 + [child.encode()]
        return tuple(info)

    pb.getJointInfo = getJointInfo

    def setJointMotorControl2(robot, idx, mode, force=0.0, **k):
        name = joints[idx][0]
        if mode == pb.TORQUE_CONTROL:
            S.tau[joint_of_name[name]] = force
        else:
            assert force == 0  # the backend only uses velocity control to switch the default motors off

    pb.setJointMotorControl2 = setJointMotorControl2

    def resetBasePositionAndOrientation(robot, position, quat_xyzw):
        st = state()
        st[_abi.ST_POS:_abi.ST_POS + 3] = position
        st[_abi.ST_QUAT:_abi.ST_QUAT + 4] = [quat_xyzw[3], quat_xyzw[0], quat_xyzw[1], quat_xyzw[2]]
        put(st)

    def resetBaseVelocity(robot, linear, angular):
        st = state()
        st[_abi.ST_LINVEL:_abi.ST_LINVEL + 3] = linear
        st[_abi.ST_ANGVEL:_abi.ST_ANGVEL + 3] = angular  # Bullet takes world-frame velocities
        put(st)

    def resetJointState(robot, idx, value, targetVelocity=0.0):
        j = joint_of_name[joints[idx][0]]
        st = state()
        st[_abi.ST_Q + j] = value
        st[_abi.ST_QD + j] = targetVelocity
        put(st)

    pb.resetBasePositionAndOrientation, pb.resetBaseVelocity, pb.resetJointState = (
        resetBasePositionAndOrientation, resetBaseVelocity, resetJointState)

    def getBasePositionAndOrientation(robot):
        st = state()
        q = st[_abi.ST_QUAT:_abi.ST_QUAT + 4]
        return tuple(st[_abi.ST_POS:_abi.ST_POS + 3]), (q[1], q[2], q[3], q[0])

    def getBaseVelocity(robot):
        st = state()
        return tuple(st[_abi.ST_LINVEL:_abi.ST_LINVEL + 3]), tuple(st[_abi.ST_ANGVEL:_abi.ST_ANGVEL + 3])

    def getJointState(robot, idx, physicsClientId=0):
        j = joint_of_name[joints[idx][0]]
        st = state()
        return (st[_abi.ST_Q + j], st[_abi.ST_QD + j], (0.0,) * 6, 0.0)

    def getLinkState(robot, idx, computeLinkVelocity=False, computeForwardKinematics=False):
        child = joints[idx][3]
        assert child == "imu"  # the only link the backend asks for (besides the base position for world-frame pushes)
        st = state()
        R = R_of(st[_abi.ST_QUAT:_abi.ST_QUAT + 4])
        p = np.asarray(b200_model.imu_position, dtype=float)
        Rbi = np.asarray(b200_model.rotation_base_to_imu, dtype=float)
        Rwi = R @ Rbi.T
        # quaternion of Rwi (x, y, z, w)
        from scipy.spatial.transform import Rotation

        q = Rotation.from_matrix(Rwi).as_quat()
        pos = st[_abi.ST_POS:_abi.ST_POS + 3] + R @ p
        om = st[_abi.ST_ANGVEL:_abi.ST_ANGVEL + 3]
        v = st[_abi.ST_LINVEL:_abi.ST_LINVEL + 3] + np.cross(om, R @ p)
        return (tuple(pos), tuple(q), (0, 0, 0), (0, 0, 0, 1), tuple(pos), tuple(q), tuple(v), tuple(om))

    pb.getBasePositionAndOrientation, pb.getBaseVelocity, pb.getJointState, pb.getLinkState = (
        getBasePositionAndOrientation, getBaseVelocity, getJointState, getLinkState)

    def getContactPoints(bodyA=None, bodyB=None, linkIndexA=None, linkIndexB=None):
        # the oracle keeps one flag: some tire touched the floor in the last collision pass; the backend ORs the wheels
        return [("contact",)] if state()[_abi.ST_CONTACT] > 0.5 else []

    pb.getContactPoints = getContactPoints

    def moving_body(idx):
        name, jtype, parent, child = joints[idx]
        return body_of_link[child] if jtype != "fixed" else None

    def getDynamicsInfo(robot, idx):
        b = moving_body(idx)
        if b is None:
            return (0.0, 0.5, (0.0, 0.0, 0.0))
        I = b200_model.inertia[b]
        return (float(b200_model.mass[b]), 0.5, (float(I[0]), float(I[1]), float(I[2])))

    def changeDynamics(robot, idx, mass=None, localInertiaDiagonal=None):
        b = moving_body(idx)
        if b is None:
            assert mass == 0.0
            return
        eps = mass / float(b200_model.mass[b]) - 1.0
        assert np.allclose(np.asarray(localInertiaDiagonal), np.asarray(b200_model.inertia[b][:3]) * (1 + eps))
        S.eps[b - 1] = eps
        S.sim.set_randomization(inertia_eps=S.eps.reshape(1, 6))

    pb.getDynamicsInfo, pb.changeDynamics = getDynamicsInfo, changeDynamics

    def applyExternalForce(robot, link, force, position, flags):
        b = 0 if link == -1 else body_of_link[joints[link][3]]
        S.ext[b] = (np.asarray(force, dtype=float), flags == pb.LINK_FRAME)

    pb.applyExternalForce = applyExternalForce

    def stepSimulation():
        if S.ext:
            f = np.zeros((1, 7, 3))
            mask = 0
            for b, (force, local) in S.ext.items():
                f[0, b] = force
                mask |= (1 << b) if local else 0
            S.sim.set_external_forces(f, mask)
        else:
            S.sim.set_external_forces(None, 0)
        S.sim.substep(S.tau.reshape(1, 6), S.h)
        S.ext = {}  # Bullet clears applied external forces and TORQUE_CONTROL torques after every step
        S.tau[:] = 0.0

    pb.stepSimulation = stepSimulation
    pb._S = S
    data = types.ModuleType("pybullet_data")
    data.getDataPath = lambda: ""
    return pb, data


def main():
    import make_wrapper_golden as wg

    wg.install_fake_gymnasium()
    wg.load_reference()

    def load(name, rel):
        import importlib.util

        spec = importlib.util.spec_from_file_location(name, os.path.join(wg.REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("upkie.utils.joystick", "upkie/utils/joystick.py")
    load("upkie.utils.point_contact", "upkie/utils/point_contact.py")
    from upkie_b200 import _abi
    from upkie_b200.model import Model as B200Model
    from upkie_b200.urdf import write_urdf

    b200_model = B200Model.standard_upkie()
    urdf_path = os.path.join(tempfile.mkdtemp(), "robot.urdf")
    write_urdf(b200_model, urdf_path, split_fixed_links=False)  # every moving link is one body of the oracle
    b200_model = B200Model.from_urdf(urdf_path)  # what the stand-in simulates is what the URDF says
    pb, data = make_fake_pybullet(b200_model, urdf_path)
    sys.modules["pybullet"] = pb
    sys.modules["pybullet_data"] = data
    backend_mod = load("upkie.envs.backends.pybullet_backend", "upkie/envs/backends/pybullet_backend.py")
    RefModel = sys.modules["upkie.model"].Model
    JointProperties = sys.modules["upkie.model"].JointProperties
    RobotState = sys.modules["upkie.utils.robot_state"].RobotState
    ExternalForce = sys.modules["upkie.utils.external_force"].ExternalForce
    from scipy.spatial.transform import Rotation

    ref_model = RefModel(urdf_path)
    rng = np.random.default_rng(20260927)
    props = {"left_wheel": JointProperties(friction=0.05), "right_knee": JointProperties(friction=0.2)}
    backend = backend_mod.PyBulletBackend(dt=0.005, gui=False, model=ref_model, joint_properties=props,
                                          inertia_variation=0.15, js_path="/nonexistent")
    S = pb._S
    out = {"generator": "tests/golden/make_backend_golden.py", "urdf": open(urdf_path).read(),
           "inertia_eps": S.eps.tolist(), "joint_friction": {"left_wheel": 0.05, "right_knee": 0.2}, "episodes": []}

    def action_rows():
        a = np.zeros((6, 6))
        a[:, 0] = [rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), np.nan, rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), np.nan]
        a[:, 1] = [0, 0, rng.uniform(-8, 8), 0, 0, rng.uniform(-8, 8)]
        a[:, 2] = rng.uniform(-0.2, 0.2, 6)
        a[:, 3] = rng.uniform(0.5, 2.0, 6)
        a[:, 4] = rng.uniform(0.5, 2.0, 6)
        a[:, 5] = [16.0, 16.0, 1.7, 16.0, 16.0, 1.7]
        return a

    for ep in range(3):
        pitch = float(rng.uniform(-0.2, 0.2))
        init = RobotState(
            position_base_in_world=np.array([0.05 * ep, 0.0, 0.58 + 0.3 * (ep == 1)]),
            orientation_base_in_world=Rotation.from_euler("ZYX", [0.3 * ep, pitch, 0.02 * ep]),
            linear_velocity_base_to_world_in_world=np.array([0.1, 0.0, -0.2 * ep]),
            angular_velocity_base_in_base=np.array([0.05, 0.3, -0.1]),
            joint_configuration=np.array([0.1, -0.2, 1.0, -0.1, 0.2, -1.0]),
        )
        q = init.orientation_base_in_world.as_quat()
        init_row = np.zeros(_abi.INIT_DIM)
        init_row[0:3] = init.position_base_in_world
        init_row[3:7] = [q[3], q[0], q[1], q[2]]
        init_row[7:10] = init.linear_velocity_base_to_world_in_world
        init_row[10:13] = init.angular_velocity_base_in_base
        init_row[13:19] = init.joint_configuration
        obs = backend.reset(init)
        episode = {"init_row": init_row.tolist(), "reset_obs": wg_jsonable(obs), "steps": []}
        for t in range(40):
            a = action_rows()
            action = {"servo": {name: {key: (float(a[j, k])) for k, key in enumerate(_abi.ACT_KEYS)}
                                for j, name in enumerate(_abi.JOINT_NAMES)}}
            if ep == 2 and t == 5:
                del action["servo"]["left_hip"]  # a joint the agent does not command: no new torque for it
            push = None
            if ep == 1 and t in (10, 11, 12):
                push = {"base": [5.0, -3.0, 0.0], "left_wheel_hub": [0.0, 2.0, 1.0]}
                backend.set_external_forces({"base": ExternalForce([5.0, -3.0, 0.0], local=False),
                                             "left_wheel_hub": ExternalForce([0.0, 2.0, 1.0], local=True)})
            if ep == 1 and t == 13:
                push = {"base": [0.0, 0.0, 0.0], "left_wheel_hub": [0.0, 0.0, 0.0]}
                backend.set_external_forces({"base": ExternalForce([0.0, 0.0, 0.0]),
                                             "left_wheel_hub": ExternalForce([0.0, 0.0, 0.0], local=True)})
            obs = backend.step(action)
            rows = [[None if v != v else v for v in r] for r in a.tolist()]
            if ep == 2 and t == 5:
                rows[0] = None
            episode["steps"].append({"action": rows, "push": push, "obs": wg_jsonable(obs)})
        out["episodes"].append(episode)
        print("episode", ep, "final pitch", obs["base_orientation"]["pitch"], "contact", obs["floor_contact"]["contact"])
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


def wg_jsonable(obs):
    def conv(x):
        if isinstance(x, dict):
            return {k: conv(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return [conv(v) for v in x]
        if hasattr(x, "tolist"):
            return x.tolist()
        if isinstance(x, (np.floating, np.bool_)):
            return x.item()
        return x

    return conv(obs)


if __name__ == "__main__":
    main()

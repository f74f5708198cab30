#!/usr/bin/env python
# SPDX-License-Identifier: Apache-2.0
"""Generate golden vectors by IMPORTING the reference's own leaf modules.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference package cannot be imported as a whole here (gymnasium, pybullet,
upkie_description, qpmpc and proxsuite are absent), but the leaf modules of the
hot path can, under a namespace stub of the ``upkie`` package:
``upkie.utils.{filters, clamp, rotations, robot_state,
robot_state_randomization}``. Their outputs on seeded inputs are written to
``tests/golden/reference_vectors.json``; the tests compare the oracle and the
host-side mirror (``upkie_b200.robot_state``) against them. Nothing at test
time reads /root/reference.
"""

import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("UPKIE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")


def load_reference_modules():
    pkg = types.ModuleType("upkie")
    pkg.__path__ = [os.path.join(REF, "upkie")]
    sys.modules["upkie"] = pkg
    utils = types.ModuleType("upkie.utils")
    utils.__path__ = [os.path.join(REF, "upkie", "utils")]
    sys.modules["upkie.utils"] = utils

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("upkie.exceptions", "upkie/exceptions.py")
    load("upkie.logging", "upkie/logging.py")
    mods = {}
    for m in ("filters", "clamp", "rotations", "robot_state_randomization", "robot_state"):
        mods[m] = load(f"upkie.utils.{m}", f"upkie/utils/{m}.py")
    return mods


def main():
    mods = load_reference_modules()
    rng = np.random.default_rng(20260923)
    out = {"generator": "tests/golden/make_golden.py", "reference_commit": "0a82a89b011cd179b20f572c486787ffc0ea69d6"}

    # upkie/utils/filters.py:63-80
    cases = []
    for _ in range(16):
        prev, new = rng.uniform(-5, 5, 2)
        dt = float(rng.choice([0.001, 0.005, 0.01]))
        tau = float(rng.choice([0.1, 0.2, 1.0]))
        cases.append({"prev": prev, "cutoff": tau, "new": new, "dt": dt,
                      "out": mods["filters"].low_pass_filter(prev, tau, new, dt)})
    out["low_pass_filter"] = cases

    # upkie/utils/clamp.py:15-61
    cases = []
    for v, lo, hi in [(0.5, 0.0, 1.0), (-0.5, 0.0, 1.0), (1.5, 0.0, 1.0), (float("nan"), 0.0, 1.0), (3.0, -2.0, 2.0),
                      (-3.0, -2.0, 2.0), (0.0, 0.0, 0.0)]:
        r = mods["clamp"].clamp(v, lo, hi)
        cases.append({"value": None if v != v else v, "lower": lo, "upper": hi, "out": None if r != r else r})
    out["clamp"] = cases
    out["clamp_abs"] = [{"value": v, "bound": b, "out": mods["clamp"].clamp_abs(v, b)}
                        for v, b in [(0.3, 1.0), (-4.0, 3.0), (4.0, 3.0)]]

    # upkie/utils/rotations.py:14-71
    quats, mats, quats_back = [], [], []
    special = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [0.5, 0.5, 0.5, 0.5], [0.1, 0.9, 0.3, -0.2],
               [0.1, -0.2, 0.9, 0.3], [0.1, 0.3, -0.2, 0.9], [-0.7, 0.1, 0.1, 0.7]]
    for k in range(40):
        q = np.array(special[k], dtype=float) if k < len(special) else rng.normal(size=4)
        q /= np.linalg.norm(q)
        R = mods["rotations"].rotation_matrix_from_quaternion(tuple(q))
        qb = mods["rotations"].quaternion_from_rotation_matrix(R)
        quats.append(q.tolist())
        mats.append(R.tolist())
        quats_back.append(np.asarray(qb).tolist())
    out["rotations"] = {"quat_wxyz": quats, "matrix": mats, "quat_from_matrix_wxyz": quats_back}

    # upkie/utils/robot_state.py:175-196 + robot_state_randomization.py:135-189
    RS = mods["robot_state"].RobotState
    RSR = mods["robot_state_randomization"].RobotStateRandomization
    from scipy.spatial.transform import Rotation

    samples = []
    for seed in range(12):
        rand = RSR(roll=0.1 * (seed % 3), pitch=0.3, x=0.05 * (seed % 2), z=0.1, omega_x=0.2, omega_y=0.5,
                   linear_velocity=np.array([0.3, 0.05, 0.1]))
        nominal_q = Rotation.from_euler("ZYX", [0.3 * (seed % 4), 0.05 * seed, 0.0])
        rs = RS(position_base_in_world=np.array([0.1 * seed, 0.0, 0.6]), orientation_base_in_world=nominal_q,
                randomization=rand)
        s = rs.sample_state(np.random.default_rng(seed))
        x, y, z, w = s.orientation_base_in_world.as_quat()
        nx, ny, nz, nw = nominal_q.as_quat()
        samples.append({
            "seed": seed,
            "randomization": {"roll": rand.roll, "pitch": rand.pitch, "x": rand.x, "z": rand.z,
                              "omega_x": rand.omega_x, "omega_y": rand.omega_y, "linear_velocity": [0.3, 0.05, 0.1]},
            "nominal_position": rs.position_base_in_world.tolist(),
            "nominal_quat_wxyz": [nw, nx, ny, nz],
            "position": s.position_base_in_world.tolist(),
            "quat_wxyz": [w, x, y, z],
            "linear_velocity": s.linear_velocity_base_to_world_in_world.tolist(),
            "angular_velocity": s.angular_velocity_base_in_base.tolist(),
        })
    out["robot_state_samples"] = samples

    # msgpack wire format: the reference's own serialize() with the Packer settings of
    # upkie/envs/backends/spine/spine_interface.py:46 (msgpack.Packer(default=serialize, use_bin_type=True))
    import msgpack

    spec = importlib.util.spec_from_file_location(
        "upkie_serialize", os.path.join(REF, "upkie/envs/backends/spine/serialize.py"))
    ser = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ser)
    packer = msgpack.Packer(default=ser.serialize, use_bin_type=True)
    joint_names = ("left_hip", "left_knee", "left_wheel", "right_hip", "right_knee", "right_wheel")
    keys = ("position", "velocity", "feedforward_torque", "kp_scale", "kd_scale", "maximum_torque")
    a = np.zeros((6, 6))
    a[:, 0] = [0.1, -0.2, np.nan, 0.3, -0.4, np.nan]
    a[:, 1] = [0.0, 0.5, -7.0, 0.25, 0.0, 7.0]
    a[:, 2] = [0.0, 0.0, 0.125, 0.0, 0.0, -0.125]
    a[:, 3] = 1.0
    a[:, 4] = [1.0, 1.0, 0.5, 1.0, 1.0, 0.5]
    a[:, 5] = [16.0, 16.0, 1.7, 16.0, 16.0, 1.7]
    action = {"servo": {name: {key: float(a[j, k]) for k, key in enumerate(keys)} for j, name in enumerate(joint_names)}}
    nested = {"imu": {"linear_acceleration": np.array([0.5, -1.5, 9.81])}, "n": 3, "flag": True, "name": "upkie"}
    out["wire"] = {"action_hex": packer.pack(action).hex(), "nested_hex": packer.pack(nested).hex()}

    # upkie.model.Model / KinematicTree (upkie/model/model.py:57-110, kinematic_tree.py:52-143) on a URDF written by
    # upkie_b200.urdf.write_urdf (the real upkie_description is absent): what the reference parses out of it
    # (wheel radius / base, wheeledness, base -> IMU rotation, joint order and limits) pins upkie_b200's own loader.
    import tempfile

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from upkie_b200.model import Model as B200Model
    from upkie_b200.urdf import write_urdf

    stub = types.ModuleType("upkie_description")
    stub.URDF_PATH = ""
    sys.modules["upkie_description"] = stub
    model_pkg = types.ModuleType("upkie.model")
    model_pkg.__path__ = [os.path.join(REF, "upkie", "model")]
    sys.modules["upkie.model"] = model_pkg

    def load_model_module(name):
        spec = importlib.util.spec_from_file_location(f"upkie.model.{name}", os.path.join(REF, "upkie", "model", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"upkie.model.{name}"] = mod
        spec.loader.exec_module(mod)
        return mod

    for name in ("se3", "joint_limit", "joint", "collision_geometry", "link", "kinematic_tree"):
        load_model_module(name)
    ref_model_mod = load_model_module("model")
    urdf_cases = []
    for split, flip in ((True, False), (False, False), (True, True)):
        m = B200Model.standard_upkie()
        if flip:  # a right-wheeled variant (Cookie-style): wheel axes reversed
            m.joint_axis = m.joint_axis.copy()
            m.joint_axis[[2, 5]] *= -1.0
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "robot.urdf")
            write_urdf(m, path, split_fixed_links=split)
            text = open(path).read()
            rm = ref_model_mod.Model(path)
            urdf_cases.append({
                "urdf": text,
                "wheel_radius": rm.wheel_radius,
                "wheel_base": rm.wheel_base,
                "left_wheeled": rm.left_wheeled,
                "rotation_base_to_imu": np.asarray(rm.rotation_base_to_imu).tolist(),
                "joint_names": [j.name for j in rm.joints],
                "joint_limits": [[j.limit.lower, j.limit.upper, j.limit.velocity, j.limit.effort] for j in rm.joints],
                "upper_leg_joints": [j.name for j in rm.upper_leg_joints],
                "wheel_joints": [j.name for j in rm.wheel_joints],
            })
    out["urdf_model"] = urdf_cases

    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

# SPDX-License-Identifier: Apache-2.0
"""Replays tests/golden/backend_runs.json: the reference's OWN PyBulletBackend class driven through reset / step /
set_external_forces / randomize_inertias on a stand-in ``pybullet`` whose physics is the oracle
(tests/golden/make_backend_golden.py). The oracle's restatement of the backend (reset quirks, moteus torque law with
friction, substep loop, observation assembly) must reproduce every spine observation; then the kernels' fp32
arithmetic is checked on the first ticks of each episode. Rows a4, a5, a7, a8, a9 of SURVEY.md section 8."""
import json
import os

import numpy as np
import pytest

from hostsim_wrap import HostSim
from upkie_b200 import _abi
from upkie_b200.model import Model

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "backend_runs.json")
A = _abi


@pytest.fixture(scope="module")
def golden():
    return json.load(open(GOLDEN))


@pytest.fixture(scope="module")
def urdf_model(golden, tmp_path_factory):
    path = tmp_path_factory.mktemp("urdf") / "robot.urdf"
    path.write_text(golden["urdf"])
    return Model.from_urdf(str(path))


def _config(golden):
    cfg = A.default_sim_config()
    cfg.skip_action_clamps = 1  # Backend.step receives spine actions as they are (no UpkieServos in front)
    for name, value in golden["joint_friction"].items():
        cfg.joint_friction[A.JOINT_NAMES.index(name)] = value
    return cfg


def _row64(obs):
    """Spine observation dictionary -> flat float64 row (layout of include/upkie_b200.h)."""
    r = np.zeros(A.SPINE_DIM)
    bo, imu = obs["base_orientation"], obs["imu"]
    r[A.SP_BASE_ANGVEL:A.SP_BASE_ANGVEL + 3] = bo["angular_velocity"]
    r[A.SP_BASE_LINVEL:A.SP_BASE_LINVEL + 3] = bo["linear_velocity"]
    r[A.SP_PITCH] = bo["pitch"]
    r[A.SP_ROT:A.SP_ROT + 9] = np.asarray(bo["rotation_base_to_world"]).reshape(9)
    r[A.SP_IMU_QUAT:A.SP_IMU_QUAT + 4] = imu["orientation"]
    r[A.SP_IMU_ANGVEL:A.SP_IMU_ANGVEL + 3] = imu["angular_velocity"]
    r[A.SP_IMU_LINACC:A.SP_IMU_LINACC + 3] = imu["linear_acceleration"]
    r[A.SP_IMU_RAWACC:A.SP_IMU_RAWACC + 3] = imu["raw_linear_acceleration"]
    r[A.SP_CONTACT] = 1.0 if obs["floor_contact"]["contact"] else 0.0
    for j, name in enumerate(A.JOINT_NAMES):
        for k, key in enumerate(A.OBS_KEYS):
            r[A.SP_SERVO + 5 * j + k] = obs["servo"][name][key]
    r[A.SP_ODOM_POS] = obs["wheel_odometry"]["position"]
    r[A.SP_ODOM_VEL] = obs["wheel_odometry"]["velocity"]
    return r


def _action(rows):
    """Golden action rows -> [6, 6] array and the joints the agent left out (no command: zero torque, the observed
    torque keeps its last value, pybullet_backend.py:276-294)."""
    a = np.zeros((6, 6))
    absent = []
    for j, r in enumerate(rows):
        if r is None:
            a[j] = [np.nan, 0.0, 0.0, 0.0, 0.0, 0.0]
            absent.append(j)
        else:
            a[j] = [np.nan if v is None else v for v in r]
    return a, absent


def _same_quaternion(a, b, tol):
    return min(np.abs(a - b).max(), np.abs(a + b).max()) < tol


def _compare(mine, ref, tol, skip_torque=()):
    cols = np.ones(A.SPINE_DIM, dtype=bool)
    cols[A.SP_IMU_QUAT:A.SP_IMU_QUAT + 4] = False
    for j in skip_torque:
        cols[A.SP_SERVO + 5 * j + 2] = False
    assert np.abs(mine[cols] - ref[cols]).max() < tol, np.argmax(np.abs(mine[cols] - ref[cols]))
    assert _same_quaternion(mine[A.SP_IMU_QUAT:A.SP_IMU_QUAT + 4], ref[A.SP_IMU_QUAT:A.SP_IMU_QUAT + 4], tol)


def test_oracle_backend_logic_matches_the_reference_class(golden, urdf_model, oracle_lib):
    cfg = _config(golden)
    osim = oracle_lib.OracleSim(urdf_model, cfg, 1, threads=1)  # one simulator across episodes, like the backend
    osim.set_randomization(inertia_eps=np.asarray(golden["inertia_eps"]).reshape(1, 6))
    assert np.abs(golden["inertia_eps"]).max() <= 0.15 and np.abs(golden["inertia_eps"]).min() > 0  # randomize_inertias ran
    saw_push = saw_contact = saw_flight = False
    for ep, episode in enumerate(golden["episodes"]):
        osim.set_external_forces(None, 0)
        osim.reset(np.asarray(episode["init_row"]).reshape(1, -1))
        # after a reset the IMU finite-difference state and the last torques of the PREVIOUS episode are still
        # there (pybullet_backend.py:157,163,220-232): the golden reset observations contain them
        _compare(osim.spine_obs()[0], _row64(episode["reset_obs"]), 1e-9)
        for t, step in enumerate(episode["steps"]):
            if step["push"] is not None:
                f = np.zeros((1, 7, 3))
                f[0, 0] = step["push"]["base"]
                f[0, urdf_model.link_body["left_wheel_hub"]] = step["push"]["left_wheel_hub"]
                osim.set_external_forces(f, 1 << urdf_model.link_body["left_wheel_hub"])
                saw_push |= bool(np.any(f))
            a, absent = _action(step["action"])
            osim.step_servos(a.reshape(1, 6, 6))
            ref = _row64(step["obs"])
            _compare(osim.spine_obs()[0], ref, 1e-8, skip_torque=absent)
            saw_contact |= ref[A.SP_CONTACT] > 0.5
            saw_flight |= ref[A.SP_CONTACT] < 0.5
    assert saw_push and saw_contact and saw_flight


def test_kernel_arithmetic_follows_the_reference_backend(golden, urdf_model):
    """fp32 kernel code (CPU build), re-synchronised on the golden state every tick is not possible (the golden
    holds observations, not states), so: same resets, same actions, compare the first ticks of each episode."""
    cfg = _config(golden)
    hs = HostSim(urdf_model, cfg, 1)
    hs.set_randomization(inertia_eps=np.asarray(golden["inertia_eps"], dtype=np.float32).reshape(1, 6))
    for episode in golden["episodes"]:
        hs.reset(np.asarray(episode["init_row"], dtype=np.float32).reshape(1, -1))
        for t, step in enumerate(episode["steps"][:8]):
            if step["push"] is not None:
                break
            a, absent = _action(step["action"])
            if absent:
                break
            hs.step_servos(a.astype(np.float32).reshape(1, 6, 6))
            mine = hs.spine_obs()[0].astype(np.float64)
            ref = _row64(step["obs"])
            cols = np.ones(A.SPINE_DIM, dtype=bool)
            cols[A.SP_IMU_QUAT:A.SP_IMU_QUAT + 4] = False
            cols[A.SP_IMU_LINACC:A.SP_IMU_RAWACC + 3] = False  # finite difference over dt: fp32 velocity noise x 200
            d = np.abs(mine - ref)
            assert d[cols].max() < 2e-2 and np.median(d[cols]) < 1e-5, (t, np.argmax(d * cols), d[cols].max())
            assert d[A.SP_IMU_LINACC:A.SP_IMU_RAWACC + 3].max() < 0.5
            assert _same_quaternion(mine[A.SP_IMU_QUAT:A.SP_IMU_QUAT + 4], ref[A.SP_IMU_QUAT:A.SP_IMU_QUAT + 4], 1e-5)

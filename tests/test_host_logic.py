# SPDX-License-Identifier: Apache-2.0
"""Host-side logic and the C-ABI surface (no GPU compute calls).

Mirrors the reference's own wrapper tests: tests/envs/test_upkie_servos.py,
test_upkie_gyropod.py, test_upkie_pendulum.py (space structure, dtypes, bounds,
neutral action), tests/utils/test_robot_state.py.
"""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from upkie_b200 import _abi
from upkie_b200.envs import (
    PENDULUM_OBS_INDICES,
    make_config,
    make_gyropod_spaces,
    make_pendulum_spaces,
    make_servo_spaces,
    servo_action_dict_to_array,
    servo_obs_array_to_dict,
    spine_row_to_dict,
)
from upkie_b200.exceptions import UpkieException, UpkieRuntimeError
from upkie_b200.robot_state import RobotState, RobotStateRandomization, quat_from_euler_zyx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- spaces ---------------------------------------------------------------------------

def test_servo_spaces_structure(model):
    act, obs, neutral, amax, amin = make_servo_spaces(model, max_gain_scale=5.0)
    assert list(act.spaces.keys()) == list(_abi.JOINT_NAMES)
    for name in _abi.JOINT_NAMES:
        assert list(act[name].spaces.keys()) == list(_abi.ACT_KEYS)  # upkie_servos.py:98-105
        assert list(obs[name].spaces.keys()) == list(_abi.OBS_KEYS)
        for key in _abi.ACT_KEYS:
            b = act[name][key]
            assert b.shape == (1,) and b.dtype == np.float32
        for key in _abi.OBS_KEYS:
            b = obs[name][key]
            assert b.shape == (1,) and b.dtype == np.float32
    hip = act["left_hip"]
    assert hip["position"].low[0] == np.float32(-1.26) and hip["position"].high[0] == np.float32(1.26)
    assert hip["velocity"].high[0] == np.float32(28.8) and hip["feedforward_torque"].high[0] == np.float32(16.0)
    assert hip["kp_scale"].low[0] == 0.0 and hip["kp_scale"].high[0] == 5.0
    assert hip["maximum_torque"].low[0] == 0.0 and hip["maximum_torque"].high[0] == 16.0
    wheel = act["right_wheel"]
    assert math.isinf(wheel["position"].high[0]) and wheel["velocity"].high[0] == np.float32(111.0)
    assert wheel["maximum_torque"].high[0] == np.float32(1.7)
    assert obs["left_knee"]["temperature"].low[0] == 0.0 and obs["left_knee"]["temperature"].high[0] == 100.0
    assert obs["left_knee"]["voltage"].low[0] == 10.0 and obs["left_knee"]["voltage"].high[0] == 44.0


def test_neutral_action(model):
    _, _, neutral, amax, amin = make_servo_spaces(model)
    for j in model.joints:
        na = neutral[j.name]  # upkie_servos.py:255-262
        assert math.isnan(na["position"]) and na["velocity"] == 0.0 and na["feedforward_torque"] == 0.0
        assert na["kp_scale"] == 1.0 and na["kd_scale"] == 1.0 and na["maximum_torque"] == j.limit.effort
        assert amin[j.name]["maximum_torque"] == 0.0 and amax[j.name]["kp_scale"] == 5.0


def test_invalid_gain_scale(model):
    for bad in (0.0, -1.0, 10.0, 12.0):
        with pytest.raises(UpkieRuntimeError):
            make_servo_spaces(model, max_gain_scale=bad)  # upkie_servos.py:148-149


def test_gyropod_and_pendulum_spaces():
    act, obs = make_gyropod_spaces(3.0, 1.0)
    assert act.shape == (2,) and act.dtype == np.float32 and obs.shape == (6,) and obs.dtype == np.float32
    assert np.array_equal(act.high, np.array([3.0, 1.0], dtype=np.float32))
    expect = np.array([np.inf, np.pi, np.inf, 3.0, 1000.0, 1.0], dtype=np.float32)
    assert np.array_equal(obs.high, expect) and np.array_equal(obs.low, -expect)
    pact, pobs = make_pendulum_spaces(2.5)
    assert pact.shape == (1,) and pact.high[0] == np.float32(2.5)
    assert PENDULUM_OBS_INDICES == [1, 0, 4, 3]
    assert np.array_equal(pobs.high, np.array([np.pi, np.inf, 1000.0, 2.5], dtype=np.float32))


def test_action_and_observation_dict_conversion(model):
    _, _, neutral, _, _ = make_servo_spaces(model)
    n = 3
    action = {
        "left_wheel": {"velocity": np.full((n, 1), 2.0, dtype=np.float32)},
        "right_hip": {"position": np.array([[0.1], [0.2], [0.3]]), "kp_scale": np.ones((n, 1))},
    }
    a = servo_action_dict_to_array(action, neutral, n)
    assert a.shape == (n, 6, 6) and a.dtype == np.float32
    assert np.all(a[:, 2, 1] == 2.0) and np.isnan(a[:, 2, 0]).all()  # missing keys -> neutral
    assert np.allclose(a[:, 3, 0], [0.1, 0.2, 0.3]) and np.all(a[:, 3, 5] == 16.0)
    assert np.isnan(a[:, 0, 0]).all() and np.all(a[:, 0, 3] == 1.0)  # missing joint -> neutral action
    obs = np.arange(n * 30, dtype=np.float32).reshape(n, 6, 5)
    d = servo_obs_array_to_dict(obs)
    assert d["left_knee"]["torque"].shape == (n, 1) and d["left_knee"]["torque"].dtype == np.float32
    assert d["right_wheel"]["voltage"][1, 0] == obs[1, 5, 4]


def test_spine_row_to_dict_keys():
    row = np.arange(_abi.SPINE_DIM, dtype=np.float32)
    d = spine_row_to_dict(row)
    # pybullet_backend.py:325-331,363-368,425-430,444-446,467-473,487-490
    assert set(d) == {"base_orientation", "floor_contact", "imu", "servo", "wheel_odometry"}
    assert set(d["base_orientation"]) == {"angular_velocity", "linear_velocity", "pitch", "rotation_base_to_world"}
    assert set(d["imu"]) == {"orientation", "angular_velocity", "linear_acceleration", "raw_linear_acceleration"}
    assert set(d["servo"]["left_hip"]) == set(_abi.OBS_KEYS)
    assert d["wheel_odometry"]["position"] == 60.0 and d["base_orientation"]["pitch"] == 6.0
    assert np.asarray(d["base_orientation"]["rotation_base_to_world"]).shape == (3, 3)
    assert isinstance(d["floor_contact"]["contact"], bool)


def test_make_config_follows_reference_defaults():
    cfg = make_config()
    assert cfg.dt == pytest.approx(1 / 200.0) and cfg.nb_substeps == 5  # int(1000 * dt), pybullet_backend.py:85-87
    assert make_config(frequency=100.0).nb_substeps == 10
    assert make_config(frequency=1000.0).nb_substeps == 1
    assert cfg.torque_control_kp == 20.0 and cfg.torque_control_kd == 1.0 and cfg.gravity == 9.81
    assert cfg.fall_pitch == 1.0 and cfg.max_ground_velocity == 3.0 and cfg.max_yaw_velocity == 1.0
    assert list(cfg.init_position) == [0.0, 0.0, 0.6]  # upkie_env.py:87-90

    class JP:
        friction = 0.25

    assert make_config(joint_properties={"left_knee": JP()}).joint_friction[1] == 0.25
    with pytest.raises(UpkieException):
        make_config(frequency=None)  # "This environment needs a loop frequency"


# ---- RobotState (tests/utils/test_robot_state.py) -----------------------------------------

def test_robot_state_defaults_and_row():
    rs = RobotState()
    assert np.array_equal(rs.position_base_in_world, [0.0, 0.0, 0.6])
    assert np.array_equal(rs.joint_configuration, np.zeros(6))
    row = rs.to_row()
    assert row.shape == (_abi.INIT_DIM,) and row[2] == 0.6 and row[3] == 1.0
    s = rs.sample_state(np.random.default_rng(0))  # no randomisation: identical state
    assert np.array_equal(s.to_row(), row)


def test_randomization_update_and_bounds():
    r = RobotStateRandomization(pitch=0.2)
    r.update(roll=0.1, v_x=0.5, v_z=0.25, omega_y=0.3)
    assert r.roll == 0.1 and r.pitch == 0.2 and np.array_equal(r.linear_velocity, [0.5, 0.0, 0.25])
    rng = np.random.default_rng(1)
    for _ in range(50):
        q = r.sample_orientation_quat(rng)
        assert abs(np.linalg.norm(q) - 1) < 1e-12
        om = r.sample_angular_velocity(rng)
        assert om[0] == 0.0 and abs(om[1]) <= 0.3 and om[2] == 0.0
        p = r.sample_position(rng)
        assert p[1] == 0.0 and p[2] == 0.0


def test_quat_from_euler_zyx_matches_scipy():
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(2)
    for _ in range(20):
        ypr = rng.uniform(-1, 1, 3)
        x, y, z, w = Rotation.from_euler("ZYX", ypr).as_quat()
        q = quat_from_euler_zyx(*ypr)
        assert np.allclose(q, [w, x, y, z], atol=1e-14)


# ---- C ABI surface ----------------------------------------------------------------------------

def _declared_symbols():
    with open(os.path.join(ROOT, "include", "upkie_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(upkie_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from upkie_b200 import _lib
    from upkie_b200.build import build

    build()  # nvcc cross-compiles without a GPU
    names = _declared_symbols()
    assert len(names) >= 24
    assert set(names) == set(_lib.SYMBOLS), "include/upkie_b200.h and upkie_b200/_lib.py disagree"
    L = C.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(L, name), f"libupkie_b200.so lacks {name}"
    assert _lib.lib().upkie_b200_abi_version() == _abi.ABI_VERSION


def test_c_default_configs_equal_python_mirrors():
    from upkie_b200 import _lib

    c = _abi.UpkieSimConfig()
    assert _lib.lib().upkie_b200_default_config(C.byref(c)) == 0
    assert _abi.struct_to_dict(c) == _abi.struct_to_dict(_abi.default_sim_config())
    m = _abi.UpkieMpcConfig()
    assert _lib.lib().upkie_b200_default_mpc_config(C.byref(m)) == 0
    assert _abi.struct_to_dict(m) == _abi.struct_to_dict(_abi.default_mpc_config())


def test_product_fails_loudly_without_gpu(model):
    """No CPU fallback: creating a handle without a CUDA device is an error."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from upkie_b200 import _lib
    from upkie_b200.sim import UpkieSim

    with pytest.raises(UpkieRuntimeError):
        UpkieSim(4, model=model)
    h = C.c_void_p()
    ms, cfg = model.to_struct(), _abi.default_sim_config()
    rc = _lib.lib().upkie_b200_create(C.byref(ms), C.byref(cfg), 4, 0, C.byref(h))
    assert rc == -2  # UPKIE_B200_ECUDA
    assert b"no CUDA device" in _lib.lib().upkie_b200_last_error()
    mh = C.c_void_p()
    mc = _abi.default_mpc_config()
    assert _lib.lib().upkie_b200_mpc_create(C.byref(mc), 4, 0, C.byref(mh)) == -2


def test_product_never_imports_the_oracle():
    """Nothing under upkie_b200/ may reference oracle/ or the host test harness."""
    pkg = os.path.join(ROOT, "upkie_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    text = fh.read()
                assert "from oracle" not in text and "import oracle" not in text, f
                assert "libupkie_oracle" not in text and "libhostsim" not in text, f


def test_model_struct_roundtrip_and_kernel_support(model):
    s = model.to_struct()
    assert list(s.parent) == [-1, 0, 1, 2, 0, 4, 5]
    assert s.wheel_radius == 0.05 and s.left_wheeled == 1
    assert [s.joint_axis[j][1] for j in range(6)] == [1, 1, 1, -1, -1, -1]
    assert sum(s.mass) == pytest.approx(5.3382, abs=1e-9)


def test_external_force_rows(model):
    """Link names -> lumped bodies, frames -> mask bits, unknown links raise as in the reference
    (pybullet_backend.py:615-619)."""
    from upkie_b200.exceptions import UpkieRuntimeError
    from upkie_b200.model import ExternalForce

    rows, mask = model.external_force_rows(
        {"torso": ExternalForce([1.0, 2.0, 3.0]), "imu": ExternalForce([1.0, 0.0, 0.0]),
         "right_wheel_tire": ExternalForce([0.0, 0.0, -5.0], local=True)}, n=3)
    assert rows.shape == (3, 7, 3) and mask == 1 << 6
    assert np.allclose(rows[:, 0], [2.0, 2.0, 3.0]) and np.allclose(rows[:, 6], [0.0, 0.0, -5.0])
    assert not rows[:, 1:6].any()
    per_env = np.arange(9.0).reshape(3, 3)
    rows, mask = model.external_force_rows({"base": ExternalForce(per_env)}, n=3)
    assert mask == 0 and np.allclose(rows[:, 0], per_env)
    with pytest.raises(UpkieRuntimeError):
        model.external_force_rows({"no_such_link": ExternalForce([0, 0, 1])}, n=1)
    with pytest.raises(UpkieRuntimeError):
        model.external_force_rows({"torso": ExternalForce([0, 0, 1]), "imu": ExternalForce([0, 0, 1], local=True)}, n=1)
    with pytest.raises(ValueError):
        ExternalForce([1.0, 2.0])


def test_numa_cpulist_parser_and_fallbacks():
    from upkie_b200 import numa

    assert numa._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert numa._parse_cpulist("") == []
    # no GPU here: unknown topology -> nothing is bound, nothing raises
    assert numa.gpu_numa_node(0) is None or isinstance(numa.gpu_numa_node(0), int)
    assert numa.bind_to_gpu_node(0) is None or isinstance(numa.bind_to_gpu_node(0), list)


def test_env_ids_and_cookie_model(tmp_path, monkeypatch):
    """``<Robot>-B200-<Action>`` ids (upkie/envs/__init__.py:24-44) and the Cookie factory's model loading
    (entry_points.py:295-336): optional ``cookie_description``, right-wheeled model out of its URDF."""
    import sys
    import types

    import upkie_b200
    from upkie_b200.model import Model
    from upkie_b200.urdf import write_urdf

    assert set(upkie_b200.ENV_IDS) == {
        f"{robot}-B200-{name}" for robot in ("Upkie", "Cookie")
        for name in ("Servos", "Gyropod", "Pendulum", "BaseVelocity")
    }
    assert upkie_b200.ENV_IDS["Cookie-B200-BaseVelocity"] == "base_velocity"
    with pytest.raises(upkie_b200.UpkieException):
        upkie_b200.make_vec("Upkie-B200-Nope", 4)
    monkeypatch.setitem(sys.modules, "cookie_description", None)  # import raises ImportError
    with pytest.raises(upkie_b200.MissingOptionalDependency):
        upkie_b200.get_cookie_model()
    right = Model.standard_upkie()
    right.joint_axis = right.joint_axis.copy()
    right.joint_axis[[2, 5]] *= -1.0
    path = str(tmp_path / "cookie.urdf")
    write_urdf(right, path)
    stub = types.ModuleType("cookie_description")
    stub.URDF_PATH = path
    monkeypatch.setitem(sys.modules, "cookie_description", stub)
    cookie = upkie_b200.get_cookie_model()
    assert cookie.left_wheeled is False and Model.standard_upkie().left_wheeled is True
    assert cookie.wheel_radius == pytest.approx(0.05) and cookie.wheel_base == pytest.approx(0.3048, abs=1e-6)


def test_header_is_c99_and_struct_sizes_match_ctypes(tmp_path):
    """include/upkie_b200.h compiles as plain C99 (the boundary is a C ABI), and every struct a caller fills has the
    size its ctypes mirror has (field-by-field agreement of the defaults is test_c_default_configs_equal_python_mirrors)."""
    import ctypes as C
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = ("UpkieModel", "UpkieSimConfig", "UpkieMpcConfig", "UpkieObserverConfig", "UpkieWheelBalancerConfig")
    src = tmp_path / "sizes.c"
    src.write_text(
        '#include <stdio.h>\n#include "upkie_b200.h"\nint main(void) {\n'
        + "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in names)
        + "  return 0;\n}\n"
    )
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
                           "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n in names:
        assert int(out[n]) == C.sizeof(getattr(_abi, n)), n

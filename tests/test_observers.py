# SPDX-License-Identifier: Apache-2.0
"""Spine observer pipeline (BaseOrientation -> FloorContact -> WheelOdometry).

The oracle is pinned by the known answers of the reference's gtests:
upkie/cpp/observers/tests/BaseOrientationTest.cpp:40-123, FloorContactTest.cpp:45-155,
WheelOdometryObserverTest.cpp:36-87, upkie/cpp/utils/tests/low_pass_filter_test.cpp;
the kernel arithmetic (fp32) is then compared with the oracle on the host and on the GPU.
"""
import ctypes as C

import numpy as np
import pytest

from upkie_b200 import _abi

A = _abi
dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)


class OracleObservers:
    def __init__(self, oracle_lib, cfg, n):
        self.L = oracle_lib.lib()
        self.L.oracle_observers_create.restype = C.c_void_p
        self.L.oracle_observers_create.argtypes = [C.POINTER(A.UpkieObserverConfig), C.c_int]
        self.L.oracle_observers_step.argtypes = [C.c_void_p, dp, dp]
        self.L.oracle_observers_reset.argtypes = [C.c_void_p]
        self.L.oracle_observers_destroy.argtypes = [C.c_void_p]
        self.L.oracle_pitch_frame_in_parent.restype = C.c_double
        self.L.oracle_pitch_frame_in_parent.argtypes = [dp]
        self.n = n
        self.cfg = cfg
        self.h = self.L.oracle_observers_create(C.byref(cfg), n)

    def step(self, spine):
        s = np.ascontiguousarray(spine, dtype=np.float64).reshape(self.n, A.SPINE_DIM)
        out = np.zeros((self.n, A.OBSV_DIM))
        self.L.oracle_observers_step(self.h, s.ctypes.data_as(dp), out.ctypes.data_as(dp))
        return out

    def reset(self):
        self.L.oracle_observers_reset(self.h)

    def pitch(self, R):
        r = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
        return self.L.oracle_pitch_frame_in_parent(r.ctypes.data_as(dp))


def spine_row(wheel_vel=(0.0, 0.0), wheel_tau=(0.0, 0.0), leg_tau=(0.0, 0.0, 0.0, 0.0), quat=(1, 0, 0, 0), gyro=(0, 0, 0)):
    r = np.zeros(A.SPINE_DIM)
    r[A.SP_IMU_QUAT:A.SP_IMU_QUAT + 4] = quat
    r[A.SP_IMU_ANGVEL:A.SP_IMU_ANGVEL + 3] = gyro
    for k, j in enumerate((2, 5)):
        r[A.SP_SERVO + j * 5 + 1] = wheel_vel[k]
        r[A.SP_SERVO + j * 5 + 2] = wheel_tau[k]
    for k, j in enumerate((0, 1, 3, 4)):
        r[A.SP_SERVO + j * 5 + 2] = leg_tau[k]
    return r


@pytest.fixture()
def cfg250(model):
    # FloorContactTest fixture (FloorContactTest.cpp:20-37): dt = 1/250, cutoff_period = 3 dt
    c = A.default_observer_config(model)
    c.dt = 1.0 / 250.0
    c.cutoff_period = 3.0 * c.dt
    return c


def test_default_observer_config_follows_spine_defaults(model):
    c = A.default_observer_config(model)
    # upkie/envs/backends/spine_backend.py:77-105,140-165
    assert c.dt == 1e-3 and c.cutoff_period == 0.2 and c.liftoff_inertia == 1e-3
    assert c.min_touchdown_acceleration == 2.0 and c.min_touchdown_torque == 0.015 and c.touchdown_inertia == 4e-3
    assert c.upper_leg_torque_threshold == 10.0
    assert list(c.signed_radius) == [0.05, -0.05]
    assert np.allclose(np.array(list(c.rotation_base_to_imu)).reshape(3, 3), np.diag([-1.0, 1.0, -1.0]))


# ---- BaseOrientationTest.cpp ----------------------------------------------------------------------

def test_pitch_known_answers(model, oracle_lib):
    o = OracleObservers(oracle_lib, A.default_observer_config(model), 1)
    phi = 0.42  # ZeroPitch: pure yaw
    Rz = np.array([[np.cos(phi), -np.sin(phi), 0], [np.sin(phi), np.cos(phi), 0], [0, 0, 1.0]])
    assert o.pitch(Rz) == 0.0
    th = 1e-3  # CloseToZero: pure rotation about y
    Ry = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    assert o.pitch(Ry) == pytest.approx(th, abs=1e-6)
    Rn = Ry.copy()
    Rn[:, 0] *= 1.0 - 1e-2  # OrientationNotNeatlyNormalized
    assert o.pitch(Rn) == pytest.approx(th, abs=1e-6)


def test_base_pitch_from_imu_known_answer(model, oracle_lib):
    # BasePitchFromIMU (BaseOrientationTest.cpp:74-89): pitch = -0.016 +- 1e-3
    c = A.default_observer_config(model)
    Rbi = [0.0, -1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0]
    for k in range(9):
        c.rotation_base_to_imu[k] = Rbi[k]
    o = OracleObservers(oracle_lib, c, 1)
    q = (0.008472769239730098, -0.9953038144146671, -0.09639792825405252, -0.002443076206500708)
    out = o.step(spine_row(quat=q))
    assert out[0, A.OBSV_PITCH] == pytest.approx(-0.016, abs=1e-3)


def test_base_orientation_neutral_values(model, oracle_lib):
    # NeutralValues (:98-122): identity quaternion, default parameters -> R = diag(-1, -1, 1), zero pitch
    o = OracleObservers(oracle_lib, A.default_observer_config(model), 1)
    out = o.step(spine_row())
    assert out[0, A.OBSV_PITCH] == 0.0
    assert np.array_equal(out[0, A.OBSV_ANGVEL:A.OBSV_ANGVEL + 3], np.zeros(3))
    assert np.allclose(out[0, A.OBSV_ROT:A.OBSV_ROT + 9].reshape(3, 3), np.diag([-1.0, -1.0, 1.0]))
    # gyro: base angular velocity = R_base_to_imu^T w_imu (BaseOrientation.h:144-148)
    out = o.step(spine_row(gyro=(0.1, 0.2, 0.3)))
    assert np.allclose(out[0, A.OBSV_ANGVEL:A.OBSV_ANGVEL + 3], [-0.1, 0.2, -0.3])


# ---- FloorContactTest.cpp ---------------------------------------------------------------------------

def test_no_torque_no_contact(oracle_lib, cfg250):
    o = OracleObservers(oracle_lib, cfg250, 1)
    out = o.step(spine_row(wheel_vel=(0.1, 0.1)))
    assert out[0, A.OBSV_CONTACT] == 0.0


def test_big_wheel_accel_torque_means_contact(oracle_lib, cfg250):
    o = OracleObservers(oracle_lib, cfg250, 1)
    vel, dt = 10.0, cfg250.dt
    o.step(spine_row(wheel_vel=(vel, vel), wheel_tau=(10.0, 10.0)))
    out = o.step(spine_row(wheel_vel=(vel + 5.0 * dt, vel + 5.0 * dt), wheel_tau=(10.0, 10.0)))
    assert out[0, A.OBSV_CONTACT] == 1.0
    assert out[0, A.OBSV_WHEEL_CONTACT] == 1.0 and out[0, A.OBSV_WHEEL_CONTACT + 1] == 1.0
    o.reset()  # what the joystick cross button triggers (reset_contact), then no contact again
    out = o.step(spine_row(wheel_vel=(0.1, 0.1)))
    assert out[0, A.OBSV_CONTACT] == 0.0


def test_leg_torque_contact(oracle_lib, cfg250):
    o = OracleObservers(oracle_lib, cfg250, 1)
    out = o.step(spine_row(leg_tau=(0.0, 0.0, 1.0, 1.0)))  # SmallLegTorqueNoContact
    assert out[0, A.OBSV_CONTACT] == 0.0
    o.reset()
    out = o.step(spine_row(leg_tau=(100.0, 100.0, 100.0, 100.0)))  # BigLegTorqueMeansContact
    assert out[0, A.OBSV_CONTACT] == 1.0
    # low-pass with cutoff 0.01 s at dt = 4 ms: 200 * 0.4 = 80 > threshold 10
    assert out[0, A.OBSV_LEG_TORQUE] == pytest.approx(200.0 * cfg250.dt / 0.01)


# ---- WheelOdometryObserverTest.cpp --------------------------------------------------------------------

def _contact_then(o, cfg, vel):
    """Bring both wheels into contact (as BigWheelAccelTorqueMeansContact), then feed `vel`."""
    o.step(spine_row(wheel_vel=(10.0, 10.0), wheel_tau=(10.0, 10.0)))
    o.step(spine_row(wheel_vel=(10.0 + 5 * cfg.dt,) * 2, wheel_tau=(10.0, 10.0)))
    return o.step(spine_row(wheel_vel=vel, wheel_tau=(10.0, 10.0)))


def test_wheel_odometry(model, oracle_lib, cfg250):
    cfg250.signed_radius[0], cfg250.signed_radius[1] = 0.5, -0.5  # the test fixture's radii
    o = OracleObservers(oracle_lib, cfg250, 1)
    out = _contact_then(o, cfg250, (1.0, -1.0))  # GoForward: velocity = radius * 1.0
    assert out[0, A.OBSV_CONTACT] == 1.0
    assert out[0, A.OBSV_ODOM_VEL] == pytest.approx(0.5 * 1.0, abs=1e-15)
    o.reset()
    out = _contact_then(o, cfg250, (1.0, 1.0))  # TurnInPlace
    assert out[0, A.OBSV_ODOM_VEL] == pytest.approx(0.0, abs=1e-15)
    o.reset()
    out = o.step(spine_row(wheel_vel=(1.0, 1.0)))  # ZeroVelocityWhenNoContact
    assert out[0, A.OBSV_ODOM_POS] == 0.0 and out[0, A.OBSV_ODOM_VEL] == 0.0
    # integration: position += velocity * dt while in contact
    o.reset()
    _contact_then(o, cfg250, (2.0, -2.0))
    p0 = o.step(spine_row(wheel_vel=(2.0, -2.0), wheel_tau=(10.0, 10.0)))[0, A.OBSV_ODOM_POS]
    p1 = o.step(spine_row(wheel_vel=(2.0, -2.0), wheel_tau=(10.0, 10.0)))[0, A.OBSV_ODOM_POS]
    assert p1 - p0 == pytest.approx(0.5 * 2.0 * cfg250.dt, abs=1e-15)


def test_low_pass_filter_guard(model):
    """low_pass_filter throws when cutoff_period <= 2 dt (upkie/cpp/utils/low_pass_filter.h:27-34):
    the library refuses such a configuration at create time."""
    from upkie_b200 import _lib

    c = A.default_observer_config(model)
    c.dt = 0.2
    h = C.c_void_p()
    rc = _lib.lib().upkie_b200_observers_create(C.byref(c), 4, 0, C.byref(h))
    assert rc in (-1, -2)  # EINVAL (with a GPU) or ECUDA (no device is checked first)


# ---- kernel arithmetic (fp32) vs oracle (fp64) on random sequences ---------------------------------------

def _random_sequence(n, T, seed):
    rng = np.random.default_rng(seed)
    seq = np.zeros((T, n, A.SPINE_DIM))
    q = rng.normal(size=(n, 4)) * [0.2, 1.0, 0.2, 0.2]  # around the upright IMU attitude (x ~ 1 in the ARS frame)
    for t in range(T):
        q += rng.normal(size=(n, 4)) * 0.01
        seq[t, :, A.SP_IMU_QUAT:A.SP_IMU_QUAT + 4] = q / np.linalg.norm(q, axis=1, keepdims=True)
        seq[t, :, A.SP_IMU_ANGVEL:A.SP_IMU_ANGVEL + 3] = rng.normal(size=(n, 3))
        servo = seq[t, :, A.SP_SERVO:A.SP_SERVO + 30].reshape(n, 6, 5)
        phase = rng.uniform(0, 6.28, (n, 1))
        servo[:, :, 1] = 5.0 * np.sin(0.05 * t + phase) + rng.normal(size=(n, 6)) * 0.2
        servo[:, :, 2] = rng.normal(size=(n, 6)) * np.array([4, 4, 0.5, 4, 4, 0.5]) * (1 + (t // 40) % 2 * 3)
    return seq


def test_host_fp32_pipeline_matches_oracle(model, oracle_lib):
    from hostsim_wrap import lib as hostlib

    n, T = 64, 300
    cfg = A.default_observer_config(model)
    o = OracleObservers(oracle_lib, cfg, n)
    L = hostlib()
    L.hostsim_observers_create.restype = C.c_void_p
    L.hostsim_observers_create.argtypes = [C.POINTER(A.UpkieObserverConfig), C.c_int]
    L.hostsim_observers_step.argtypes = [C.c_void_p, fp, fp]
    h = L.hostsim_observers_create(C.byref(cfg), n)
    seq = _random_sequence(n, T, 0)
    flips = 0
    seen = set()
    for t in range(T):
        s32 = seq[t].astype(np.float32)
        oo = o.step(s32.astype(np.float64))
        seen.update(np.unique(oo[:, A.OBSV_CONTACT] > 0).tolist())
        go = np.zeros((n, A.OBSV_DIM), dtype=np.float32)
        L.hostsim_observers_step(h, s32.ctypes.data_as(fp), go.ctypes.data_as(fp))
        assert np.abs(go[:, A.OBSV_PITCH] - oo[:, A.OBSV_PITCH]).max() < 2e-3  # acos near 0 amplifies fp32 round-off
        assert np.abs(go[:, A.OBSV_ROT:A.OBSV_ROT + 9] - oo[:, A.OBSV_ROT:A.OBSV_ROT + 9]).max() < 1e-6
        assert np.abs(go[:, A.OBSV_LEG_TORQUE] - oo[:, A.OBSV_LEG_TORQUE]).max() < 1e-4
        flips += int((go[:, A.OBSV_CONTACT] != oo[:, A.OBSV_CONTACT]).sum())
    assert flips <= 3  # hysteresis thresholds crossed within fp32 round-off
    assert np.abs(go[:, A.OBSV_ODOM_POS] - oo[:, A.OBSV_ODOM_POS]).max() < 5e-3
    assert seen == {True, False}  # the sequence exercises both outcomes of the contact estimator


@pytest.mark.gpu
def test_gpu_pipeline_matches_oracle_and_runs_on_the_simulation(model, oracle_lib):
    import torch

    from upkie_b200.observers import ObserverPipeline
    from upkie_b200.sim import UpkieSim

    n, T = 1024, 200
    cfg = A.default_observer_config(model)
    o = OracleObservers(oracle_lib, cfg, n)
    pipe = ObserverPipeline(n, model=model)
    seq = _random_sequence(n, T, 1)
    flips = 0
    for t in range(T):
        s32 = seq[t].astype(np.float32)
        oo = o.step(s32.astype(np.float64))
        go = pipe.step(torch.from_numpy(s32).cuda()).cpu().numpy()
        assert np.abs(go[:, A.OBSV_ROT:A.OBSV_ROT + 9] - oo[:, A.OBSV_ROT:A.OBSV_ROT + 9]).max() < 1e-6
        assert np.abs(go[:, A.OBSV_PITCH] - oo[:, A.OBSV_PITCH]).max() < 2e-3
        flips += int((go[:, A.OBSV_CONTACT] != oo[:, A.OBSV_CONTACT]).sum())
    assert flips <= 20
    assert np.abs(go[:, A.OBSV_ODOM_POS] - oo[:, A.OBSV_ODOM_POS]).max() < 5e-3
    d = ObserverPipeline.row_to_dict(go[0])
    assert set(d) == {"base_orientation", "floor_contact", "wheel_odometry"}
    # on the simulation: 1 kHz "spine" (one substep per step), standing robots end up in estimated contact,
    # observer pitch agrees with the simulator's base pitch and odometry with the wheel-angle odometry
    sim_cfg = A.default_sim_config(frequency=1000.0)
    sim = UpkieSim(n, model=model, config=sim_cfg)
    sim.reset(seed=1)
    pipe.reset()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    a = torch.zeros((n, 6, 6), device="cuda")
    a[:, :, 3] = a[:, :, 4] = 1.0
    a[:, :, 5] = torch.tensor(model.tau_max, dtype=torch.float32, device="cuda")
    a[:, 2, 0] = a[:, 5, 0] = float("nan")
    for t in range(600):
        a[:, 2, 1] = 8.0 * np.sin(0.02 * t)  # drive the wheels back and forth
        a[:, 5, 1] = -8.0 * np.sin(0.02 * t)
        sim.step_servos(a)
        out = pipe.step(sim.spine_obs())
    sp = sim.spine_obs()
    assert (out[:, A.OBSV_CONTACT] > 0.5).float().mean().item() > 0.9
    assert (out[:, A.OBSV_PITCH] - sp[:, A.SP_PITCH]).abs().max().item() < 5e-3
    assert (out[:, A.OBSV_ODOM_VEL] - sp[:, A.SP_ODOM_VEL]).abs().median().item() < 1e-3

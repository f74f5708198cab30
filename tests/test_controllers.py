# SPDX-License-Identifier: Apache-2.0
"""Spine controller pipeline "wheel_balancer" (WheelStopper -> WheelBalancer).

The reference ships no numerical test for these controllers (upkie/cpp/controllers/tests/ControllerTest.cpp only
checks that the base class does nothing), so the fp64 oracle is pinned on hand-computed cycles of
WheelBalancer.cpp:35-110; the kernel arithmetic (fp32) is then compared with the oracle on the host and on the GPU,
and the closed loop (simulator -> observers -> controller -> simulator) must keep the robots upright.
"""
import ctypes as C

import numpy as np
import pytest

from upkie_b200 import _abi

A = _abi
dp = C.POINTER(C.c_double)


class OracleBalancer:
    def __init__(self, oracle_lib, cfg, n):
        self.L = oracle_lib.lib()
        self.L.oracle_wheel_balancer_create.restype = C.c_void_p
        self.L.oracle_wheel_balancer_create.argtypes = [C.POINTER(A.UpkieWheelBalancerConfig), C.c_int]
        self.L.oracle_wheel_balancer_step.argtypes = [C.c_void_p, dp, dp, dp, dp]
        self.L.oracle_wheel_balancer_destroy.argtypes = [C.c_void_p]
        self.n = n
        self.h = self.L.oracle_wheel_balancer_create(C.byref(cfg), n)

    def step(self, obs3, target, action):
        o = np.ascontiguousarray(obs3, dtype=np.float64)
        t = np.ascontiguousarray(target, dtype=np.float64) if target is not None else None
        a = np.ascontiguousarray(action, dtype=np.float64).copy()
        st = np.zeros((self.n, 4))
        self.L.oracle_wheel_balancer_step(self.h, o.ctypes.data_as(dp), t.ctypes.data_as(dp) if t is not None else None,
                                          a.ctypes.data_as(dp), st.ctypes.data_as(dp))
        return a, st

    def __del__(self):
        self.L.oracle_wheel_balancer_destroy(self.h)


def _action(n, rng=None):
    a = np.zeros((n, 6, 6))
    a[:, :, 3:5] = 1.0
    a[:, :, 5] = 1.0
    if rng is not None:
        a[:, :, 0:3] = rng.uniform(-1, 1, (n, 6, 3))
    return a


def test_oracle_known_cycles(oracle_lib):
    cfg = A.default_wheel_balancer_config(1000.0)
    ob = OracleBalancer(oracle_lib, cfg, 1)
    # cycle 1, in contact: pitch 0.1 rad, odometry 0.02 m, target velocity 0.5 m/s, yaw 0.2 rad/s
    a, st = ob.step([[0.1, 1.0, 0.02]], [[0.5, 0.2]], _action(1))
    e0, e1 = 0.0 - 0.02, 0.0 - 0.1
    integral = (1.6 * e0 + 20.0 * e1) * 1e-3
    target_pos = 0.5 * 1e-3
    v = -0.5 - (0.7 * e0 + 1.8 * e1) - integral
    assert np.allclose(st[0], [v, integral, target_pos, 0.2], rtol=0, atol=1e-15)
    ref = v / 0.06
    y2w = 0.1524 / 0.06
    assert np.isclose(a[0, 2, 1], ref + y2w * 0.2) and np.isclose(a[0, 5, 1], -ref + y2w * 0.2)
    assert np.isnan(a[0, 2, 0]) and np.isnan(a[0, 5, 0]) and a[0, 2, 2] == 0.0 and a[0, 5, 2] == 0.0
    assert (a[0, [0, 1, 3, 4], 3] == 4.0).all() and (a[0, [0, 1, 3, 4], 4] == 4.0).all()  # turning: 2 + 2
    assert (a[0, [2, 5], 3] == 1.0).all()  # wheel gains untouched
    # cycle 2, in the air: integrator and target decay towards 0 / the measured position with a 1 s period
    a, st2 = ob.step([[0.0, 0.0, 0.03]], [[0.0, 0.0]], _action(1))
    assert np.isclose(st2[0, 1], integral * (1 - 1e-3)) and np.isclose(st2[0, 2], target_pos + 1e-3 * (0.03 - target_pos))
    assert (a[0, [0, 1, 3, 4], 3] == 2.0).all()  # not turning
    # the command uses the error computed BEFORE the target update
    assert np.isclose(st2[0, 0], -(0.7 * (target_pos - 0.03)) - st2[0, 1])
    # fallen: zero velocity, state frozen; saturation of the command at max_ground_velocity
    _, st3 = ob.step([[1.2, 1.0, 0.03]], [[1.0, 0.0]], _action(1))
    assert st3[0, 0] == 0.0 and np.array_equal(st3[0, 1:3], st2[0, 1:3])
    _, st4 = ob.step([[-0.9, 1.0, 10.0]], [[0.0, 0.0]], _action(1))
    assert st4[0, 0] == 2.0  # -(0.7 * (-9.97) + 1.8 * 0.9) - integral > max_ground_velocity
    # the target position never trails the measured one by more than 1 m
    assert np.isclose(st4[0, 2], 9.0)


def test_kernel_arithmetic_matches_oracle(oracle_lib):
    from hostsim_wrap import wheel_balancer_step

    n = 512
    cfg = A.default_wheel_balancer_config(200.0)
    ob = OracleBalancer(oracle_lib, cfg, n)
    rng = np.random.default_rng(7)
    state = np.zeros((n, 4), dtype=np.float32)
    for k in range(200):
        obs3 = np.stack([rng.uniform(-1.2, 1.2, n), (rng.uniform(0, 1, n) > 0.2).astype(float), rng.uniform(-2, 2, n)], axis=1)
        target = np.stack([rng.uniform(-1, 1, n), rng.uniform(-0.3, 0.3, n)], axis=1)
        act = _action(n, rng)
        a32 = act.astype(np.float32)
        wheel_balancer_step(cfg, state, obs3, target, a32)
        a64, st = ob.step(obs3.astype(np.float32), target.astype(np.float32), act.astype(np.float32))
        assert np.allclose(state, st, rtol=1e-5, atol=2e-5)
        assert np.allclose(a32, a64, rtol=1e-5, atol=5e-4, equal_nan=True)
        assert np.array_equal(np.isnan(a32), np.isnan(a64))


@pytest.mark.gpu
def test_gpu_pipeline_matches_oracle_and_balances(model, oracle_lib):
    import torch

    from upkie_b200.controllers import WheelBalancerPipeline
    from upkie_b200.observers import ObserverPipeline
    from upkie_b200.sim import UpkieSim

    # 1. same inputs, same outputs as the oracle (observer-row layout and spine-row layout)
    n = 2048
    cfg = A.default_wheel_balancer_config(200.0)
    rng = np.random.default_rng(11)
    for width, (pi, ci, oi) in ((A.OBSV_DIM, (A.OBSV_PITCH, A.OBSV_CONTACT, A.OBSV_ODOM_POS)),
                                (A.SPINE_DIM, (A.SP_PITCH, A.SP_CONTACT, A.SP_ODOM_POS))):
        wb = WheelBalancerPipeline(n, config=cfg)
        ob = OracleBalancer(oracle_lib, cfg, n)
        for k in range(20):
            rows = np.zeros((n, width), dtype=np.float32)
            rows[:, pi] = rng.uniform(-1.2, 1.2, n)
            rows[:, ci] = rng.uniform(0, 1, n) > 0.2
            rows[:, oi] = rng.uniform(-2, 2, n)
            target = np.stack([rng.uniform(-1, 1, n), rng.uniform(-0.3, 0.3, n)], axis=1).astype(np.float32)
            act = _action(n, rng).astype(np.float32)
            out = wb.step(torch.from_numpy(rows).cuda(), torch.from_numpy(act).cuda(), torch.from_numpy(target).cuda())
            a64, st = ob.step(rows[:, [pi, ci, oi]], target, act)
            assert np.allclose(wb.state().cpu().numpy(), st, rtol=1e-5, atol=2e-5)
            assert np.allclose(out.cpu().numpy(), a64, rtol=1e-5, atol=5e-4, equal_nan=True)
        wb.reset()
        assert not wb.state().cpu().numpy().any()

    # 2. closed loop at 250 Hz: simulator -> spine observation -> observers -> wheel_balancer -> simulator.
    # The reference runs the pipeline at the spine frequency (1 kHz); here one cycle per 4 ms env tick. (At the
    # default 200 Hz the FloorContact low-pass, cutoff 0.01 s, violates cutoff > 2 dt and the reference throws.)
    n = 256
    sim = UpkieSim(n, model=model, config=A.default_sim_config(250.0))
    init = np.zeros((n, A.INIT_DIM), dtype=np.float32)
    init[:, 2] = 0.58
    pitch0 = np.random.default_rng(3).uniform(-0.15, 0.15, n)
    init[:, 3], init[:, 5] = np.cos(pitch0 / 2), np.sin(pitch0 / 2)
    sim.reset(init_state=torch.from_numpy(init).cuda())
    obs_pipe = ObserverPipeline(n, model=model, spine_frequency=250.0)
    wbc = A.default_wheel_balancer_config(250.0)
    wbc.wheel_radius = float(model.wheel_radius)
    wb = WheelBalancerPipeline(n, config=wbc)
    neutral = torch.zeros((n, 6, 6), device="cuda")
    neutral[:, :, 3:5] = 1.0
    neutral[:, :, 5] = torch.tensor(model.tau_max, device="cuda", dtype=torch.float32)
    pitches = []
    for k in range(750):  # 3 s
        spine = sim.spine_obs()
        rows = obs_pipe.step(spine)
        act = wb.step(rows, neutral.clone())
        sim.step_servos(act)
        pitches.append(rows[:, A.OBSV_PITCH].abs().max().item())
    st = sim.get_state().cpu().numpy()
    assert max(pitches[250:]) < 0.3, max(pitches[250:])  # every robot upright after the transient
    assert (st[:, A.ST_POS + 2] > 0.4).all()

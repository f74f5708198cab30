# SPDX-License-Identifier: Apache-2.0
"""GPU parity: sm_100a kernels (through the C ABI) vs the fp64 CPU oracle."""
import numpy as np
import pytest

from conftest import random_servo_actions, random_states

pytestmark = pytest.mark.gpu

# fp32 kernels vs fp64 oracle after one 5 ms tick from identical states.
# Velocities carry the touchdown sensitivity of the Bullet-style contact row
# (d v / d penetration = 1/h = 1000 1/s, times 1/r = 20 on wheel rates), so the
# worst-case velocity tolerance is looser than the position tolerance.
TOL_POS = 2e-5
TOL_VEL_STEADY = 5e-4
TOL_VEL_WORST = 2e-2


def _mk(n, model, cfg=None):
    import torch

    from upkie_b200 import _abi
    from upkie_b200.sim import UpkieSim

    cfg = cfg if cfg is not None else _abi.default_sim_config()
    return UpkieSim(n, model=model, config=cfg), cfg, torch


def test_one_tick_servos_matches_oracle(model, oracle_lib):
    n = 2048
    sim, cfg, torch = _mk(n, model)
    st = random_states(n, seed=3)
    act = random_servo_actions(n, model, seed=4)
    osim = oracle_lib.OracleSim(model, cfg, n, threads=8)
    st32 = st.astype(np.float32)
    osim.set_state(st32.astype(np.float64))
    sim.set_state(torch.from_numpy(st32).cuda())
    act32 = act.astype(np.float32)
    oobs, orew, oterm, otrunc = osim.step_servos(act32.astype(np.float64))
    gobs, grew, gterm, gtrunc = sim.step_servos(torch.from_numpy(act32).cuda())
    torch.cuda.synchronize()
    gs = sim.get_state().cpu().numpy().astype(np.float64)
    os_ = osim.get_state()
    dpos = np.abs(gs[:, :7] - os_[:, :7]).max()
    dq = np.abs(gs[:, 13:19] - os_[:, 13:19]).max()
    dvel = np.abs(gs[:, 7:13] - os_[:, 7:13])
    dqd = np.abs(gs[:, 19:25] - os_[:, 19:25])
    assert dpos < TOL_POS and dq < 2e-4, (dpos, dq)
    assert np.median(dvel.max(axis=1)) < TOL_VEL_STEADY
    assert np.median(dqd.max(axis=1)) < TOL_VEL_STEADY
    assert dvel.max() < TOL_VEL_WORST and dqd.max() < 20 * TOL_VEL_WORST, (dvel.max(), dqd.max())
    # observations: positions/velocities/torques as the state, constants exact
    g = gobs.cpu().numpy().astype(np.float64)
    assert np.array_equal(g[:, :, 3], np.full((n, 6), 42.0))
    assert np.array_equal(g[:, :, 4], np.full((n, 6), 18.0))
    assert np.abs(g[:, :, 0] - oobs[:, :, 0]).max() < 2e-4
    assert np.median(np.abs(g[:, :, 2] - oobs[:, :, 2])) < 1e-3
    # integer outputs bit-exact
    assert np.array_equal(gterm.cpu().numpy(), oterm)
    assert np.array_equal(gtrunc.cpu().numpy(), otrunc)
    assert np.array_equal(grew.cpu().numpy(), orew.astype(np.float32))
    # error flags (clamped / NaN velocity) bit-exact
    assert np.array_equal(sim.error_flags().cpu().numpy().astype(np.uint32), osim.error_flags())
    assert np.array_equal(gs[:, 40], os_[:, 40])  # floor contact flag


def test_spine_observation_matches_oracle(model, oracle_lib):
    n = 512
    sim, cfg, torch = _mk(n, model)
    st32 = random_states(n, seed=7).astype(np.float32)
    act32 = random_servo_actions(n, model, seed=8).astype(np.float32)
    osim = oracle_lib.OracleSim(model, cfg, n, threads=8)
    osim.set_state(st32.astype(np.float64))
    sim.set_state(torch.from_numpy(st32).cuda())
    osim.step_servos(act32.astype(np.float64))
    sim.step_servos(torch.from_numpy(act32).cuda())
    g = sim.spine_obs().cpu().numpy().astype(np.float64)
    o = osim.spine_obs()
    # quaternion sign convention must match exactly (scipy branch choice)
    assert np.abs(g[:, 16:20] - o[:, 16:20]).max() < 1e-4
    assert np.abs(g[:, 6] - o[:, 6]).max() < 1e-5  # pitch
    assert np.abs(g[:, 7:16] - o[:, 7:16]).max() < 1e-5  # rotation matrix
    assert np.array_equal(g[:, 29], o[:, 29])  # contact
    assert np.median(np.abs(g[:, 23:29] - o[:, 23:29])) < 5e-2  # IMU accelerations (finite differences / dt)
    assert np.abs(g[:, 60] - o[:, 60]).max() < 1e-5  # odometry position

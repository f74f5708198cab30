# SPDX-License-Identifier: Apache-2.0
"""Replays tests/golden/wrapper_trajectories.json: trajectories produced in the build container by the reference's
OWN UpkieServos / UpkieGyropod / UpkiePendulum classes running on top of the fp64 oracle
(tests/golden/make_wrapper_golden.py). Between the agent's action and the physics, and between the spine
observation and the env observation, the golden side is the reference's code; here the same seeds and actions go
through (1) the oracle's restatement of those wrappers and (2) the kernels' arithmetic compiled for the CPU.
Rows a1-a3, a10, a11 of SURVEY.md section 8."""
import json
import os

import numpy as np
import pytest

from hostsim_wrap import HostSim
from upkie_b200 import _abi
from upkie_b200.robot_state import RobotState, RobotStateRandomization

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "wrapper_trajectories.json")


@pytest.fixture(scope="module")
def cases():
    return json.load(open(GOLDEN))["cases"]


def _action_array(rows):
    return np.array([[np.nan if v is None else v for v in r] for r in rows], dtype=np.float64)


def _reset_like_the_reference_env(osim, case):
    """What UpkieEnv.reset(seed) did in the generator: sample with default_rng(seed), backend.reset(sample)."""
    row = _init_state().sample_state(np.random.default_rng(case["seed"])).to_row()
    osim.reset(row.reshape(1, -1))


def _init_state():
    return RobotState(position_base_in_world=np.array([0.0, 0.0, 0.58]),
                      randomization=RobotStateRandomization(pitch=0.05, x=0.1, omega_y=0.1))


def test_reset_sampling_through_the_reference_env(cases, model, oracle_lib):
    """UpkieEnv.reset(seed) -> init_state.sample_state(self.np_random) -> backend.reset (upkie_env.py:180-191): our
    mirror of the sampler, seeded the same way, puts the oracle in the same state the reference env did."""
    for case in cases:
        cfg = _abi.default_sim_config()
        osim = oracle_lib.OracleSim(model, cfg, 1, threads=1)
        row = _init_state().sample_state(np.random.default_rng(case["seed"])).to_row()
        osim.reset(row.reshape(1, -1))
        ref_row = np.asarray(case["state_after_reset"])  # recorded before the friction impulses joined the row (46 words)
        assert np.allclose(osim.get_state()[0][:len(ref_row)], ref_row, rtol=0, atol=1e-12), case["kind"]


@pytest.mark.parametrize("index", [0, 1, 2])
def test_gyropod_and_pendulum_wrappers_match_the_reference_classes(cases, model, oracle_lib, index):
    case = cases[index]
    act_dim = 2 if case["kind"] == "gyropod" else 1
    cfg = _abi.default_sim_config()
    osim = oracle_lib.OracleSim(model, cfg, 1, threads=1)
    _reset_like_the_reference_env(osim, case)
    assert np.allclose(osim.reset_obs(6 if act_dim == 2 else 4)[0], case["reset_obs"], rtol=0, atol=1e-7)
    for t, (a, o, term) in enumerate(zip(case["actions"], case["obs"], case["terminated"])):
        obs, rew, te, tr = osim.step_gyropod(np.asarray(a, dtype=np.float64).reshape(1, act_dim), act_dim)
        # same physics (the oracle) on both sides: what is compared is the wrapper arithmetic; observations are
        # float32 on both sides (upkie_gyropod.py:206-214)
        assert np.allclose(obs[0], o, rtol=0, atol=2e-6), (case["kind"], t, obs[0], o)
        assert bool(te[0]) == term and rew[0] == 0.0 and tr[0] == 0, (case["kind"], t)
    if index == 2:
        assert any(case["terminated"]) and not case["terminated"][0]  # the fall case really falls


def test_servos_clamping_and_dictionaries_match_the_reference_class(cases, model, oracle_lib):
    case = cases[3]
    assert case["kind"] == "servos"
    cfg = _abi.default_sim_config()  # clamps on: the oracle restates UpkieServos.get_spine_action (upkie_servos.py:308-344)
    osim = oracle_lib.OracleSim(model, cfg, 1, threads=1)
    _reset_like_the_reference_env(osim, case)
    assert np.allclose(osim.reset_obs(30)[0], case["reset_obs"], rtol=0, atol=1e-6)
    sent = [_action_array(a) for a in case["spine_actions"]]
    clamped_some = False
    for t, (a, o) in enumerate(zip(case["actions"], case["obs"])):
        raw = _action_array(a)
        clamped_some |= not np.array_equal(np.nan_to_num(raw, nan=7.0), np.nan_to_num(sent[t], nan=7.0))
        obs, _, te, _ = osim.step_servos(raw.reshape(1, 6, 6))
        assert np.allclose(obs[0], np.asarray(o), rtol=0, atol=2e-5), (t, np.abs(obs[0] - np.asarray(o)).max())
        assert te[0] == 0
    assert clamped_some  # the action sequence exercises the clamps
    assert (osim.error_flags()[0] & _abi.ERR_CLAMPED) != 0


@pytest.mark.parametrize("index", [0, 1, 2])
def test_kernel_arithmetic_follows_the_reference_wrappers(cases, model, index):
    """fp32 kernel code (CPU build) on the same seeds / actions: round-off grows along a closed-loop trajectory
    (contact rows: 1/h sensitivity), so the comparison is tight on the first ticks and statistical afterwards."""
    case = cases[index]
    act_dim = 2 if case["kind"] == "gyropod" else 1
    cfg = _abi.default_sim_config()
    hs = HostSim(model, cfg, 1)
    row = np.zeros((1, _abi.STATE_DIM), dtype=np.float32)
    row[0, :len(case["state_after_reset"])] = np.asarray(case["state_after_reset"], dtype=np.float32)
    hs.set_state(row)
    worst = 0.0
    fell_at = None
    for t, (a, o, term) in enumerate(zip(case["actions"], case["obs"], case["terminated"])):
        obs6, te = hs.step_gyropod(np.asarray(a, dtype=np.float32).reshape(1, act_dim), act_dim)
        mine = obs6[0] if act_dim == 2 else obs6[0][[1, 0, 4, 3]]
        err = np.abs(mine - np.asarray(o))
        if t < 10:
            assert err.max() < 5e-3, (case["kind"], t, err)
        worst = max(worst, err[[0, 1]].max() if act_dim == 1 else err[[0, 1, 2]].max())  # positions / angles
        if te[0] and fell_at is None:
            fell_at = t
    assert worst < 5e-2, worst
    golden_fall = next((t for t, x in enumerate(case["terminated"]) if x), None)
    assert (fell_at is None) == (golden_fall is None)
    if golden_fall is not None:
        assert abs(fell_at - golden_fall) <= 1


def test_spaces_and_neutral_action_equal_the_reference_classes(model):
    """single_action_space / single_observation_space of B200VectorEnv against the spaces the reference's own
    UpkieServos / UpkieGyropod / UpkiePendulum instances reported in the build container (bounds, shapes, dtypes)."""
    from upkie_b200.envs import make_gyropod_spaces, make_pendulum_spaces, make_servo_spaces

    g = json.load(open(GOLDEN))["spaces"]

    def same(box, ref):
        assert list(box.shape) == ref["shape"] and str(np.dtype(box.dtype)) == ref["dtype"]
        assert np.array_equal(np.asarray(box.low, dtype=float), np.asarray(ref["low"]))
        assert np.array_equal(np.asarray(box.high, dtype=float), np.asarray(ref["high"]))

    act, obs, neutral, _, _ = make_servo_spaces(model)
    assert list(act.spaces if hasattr(act, "spaces") else act) == list(g["servos_action"])  # joint order
    for joint in g["servos_action"]:
        assert list(act[joint].spaces if hasattr(act[joint], "spaces") else act[joint]) == list(g["servos_action"][joint])
        for key, ref in g["servos_action"][joint].items():
            same(act[joint][key], ref)
        for key, ref in g["servos_observation"][joint].items():
            same(obs[joint][key], ref)
        for key, ref in g["neutral_action"][joint].items():
            mine = float(neutral[joint][key])
            assert (mine != mine) if ref is None else (mine == ref), (joint, key)
    ga, go = make_gyropod_spaces()
    same(ga, g["gyropod_action"])
    same(go, g["gyropod_observation"])
    pa, po = make_pendulum_spaces()
    same(pa, g["pendulum_action"])
    same(po, g["pendulum_observation"])

# SPDX-License-Identifier: Apache-2.0
"""Replays tests/golden/base_velocity_run.json: a 300-tick run of the reference's OWN UpkieBaseVelocity env (with its
gyropod / servos wrappers and MPCBalancer shell) on top of the oracle (tests/golden/make_base_velocity_golden.py).
The replay goes through upkie_b200.base_velocity.base_velocity_tick -- the very function B200VectorEnv runs on the
GPU -- with oracle-backed callables on CPU tensors: which spine observation the MPC sees, what the gyropod is given,
and the dead reckoning along the post-step yaw (row a12 of SURVEY.md section 8)."""
import json
import os

import numpy as np
import torch

from upkie_b200 import _abi
from upkie_b200.base_velocity import base_velocity_tick, mpc_inputs_from_spine

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "base_velocity_run.json")


def test_base_velocity_tick_matches_the_reference_env(model, oracle_lib):
    run = json.load(open(GOLDEN))
    cfg = _abi.default_sim_config()
    osim = oracle_lib.OracleSim(model, cfg, 1, threads=1)
    osim.reset(np.asarray(run["init_row"]).reshape(1, -1))
    om = oracle_lib.OracleMpc(_abi.default_mpc_config())  # horizon 50: MPCBalancer defaults
    state = {"v": np.zeros(1)}

    def mpc_step_spine(target, spine, dt):
        x0, contact = mpc_inputs_from_spine(spine)
        vc, _, found, _ = om.step(x0.numpy().astype(np.float64), target.numpy().astype(np.float64), contact.numpy(), dt, state["v"])
        assert found.all()
        state["v"] = vc
        return torch.from_numpy(vc.astype(np.float32))  # the gyropod action is float32 (upkie_base_velocity.py:186-188)

    def step_gyropod(a):
        obs, rew, term, trunc = osim.step_gyropod(a.numpy().astype(np.float64), 2)
        return (torch.from_numpy(obs.astype(np.float32)), torch.from_numpy(rew.astype(np.float32)),
                torch.from_numpy(term), torch.from_numpy(trunc))

    def spine_obs():
        return torch.from_numpy(osim.spine_obs().astype(np.float64))

    spine = spine_obs()  # what reset() returned
    xy = torch.zeros((1, 2), dtype=torch.float64)
    worst = 0.0
    for t, (a, o, term, v) in enumerate(zip(run["actions"], run["obs"], run["terminated"], run["commanded_velocity"])):
        action = torch.tensor([a], dtype=torch.float32)
        obs, rew, te, tr, spine = base_velocity_tick(action, spine, xy, cfg.dt, mpc_step_spine, step_gyropod, spine_obs)
        assert abs(float(state["v"][0]) - v) < 1e-6, (t, state["v"][0], v)  # the MPC saw the same (previous) observation
        worst = max(worst, float(np.abs(obs.numpy()[0] - np.asarray(o)).max()))
        assert bool(te[0]) == term and float(rew[0]) == 0.0
    assert worst < 5e-6, worst  # x, y, yaw after 300 ticks of closed loop
    assert abs(run["obs"][-1][0]) > 0.05 and abs(run["obs"][-1][2]) > 0.05  # the run actually drives and turns

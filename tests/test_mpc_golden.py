# SPDX-License-Identifier: Apache-2.0
"""Replays tests/golden/mpc_balancer_runs.json: closed-loop runs of the reference's OWN MPCBalancer.step shell
(observation unpacking, get_target_states, post-processing; mpc_balancer.py:18-37,237-312) on stand-ins for qpmpc /
proxsuite (tests/golden/make_mpc_golden.py). Pins the oracle's independent C++ restatement (oracle_mpc_step) and, with
ProxQP's tolerance, the kernel's fp32 Riccati / active-set arithmetic."""
import json
import os

import numpy as np
import pytest

from hostsim_wrap import mpc_step
from upkie_b200 import _abi

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "mpc_balancer_runs.json")


@pytest.mark.parametrize("index", [0, 1])
def test_mpc_balancer_shell_matches_the_reference(oracle_lib, index):
    run = json.load(open(GOLDEN))["runs"][index]
    cfg = _abi.default_mpc_config()
    cfg.nb_timesteps = run["nb_timesteps"]
    om = oracle_lib.OracleMpc(cfg)
    v_oracle, v_kernel = 0.0, 0.0
    saw_fall = saw_liftoff = saw_clamped_plan = False
    for k, s in enumerate(run["steps"]):
        x0 = np.asarray(s["x0"], dtype=np.float64).reshape(1, 4)
        vt = np.array([s["target"]])
        contact = np.array([1 if s["contact"] else 0], dtype=np.uint8)
        vc, first, found, plan = om.step(x0, vt, contact, run["dt"], np.array([v_oracle]))
        v_oracle = float(vc[0])
        # fp64 oracle (own QP build + own active-set solver) vs reference shell on an exact BVLS solve
        assert abs(v_oracle - s["commanded_velocity"]) < 1e-7, (k, v_oracle, s["commanded_velocity"])
        fallen = abs(s["x0"][1]) > 1.0
        if not fallen and s["contact"]:
            assert abs(float(first[0]) - s["first_input"]) < 1e-6, (k, first[0], s["first_input"])
            saw_clamped_plan |= abs(s["first_input"]) >= 10.0 - 1e-9
        saw_fall |= fallen
        saw_liftoff |= not s["contact"]
        # what the kernel runs (fp32), one tick from the reference's previous command
        prev = run["steps"][k - 1]["commanded_velocity"] if k else 0.0
        vk, plan32, found32, _ = mpc_step(cfg, x0, vt, contact, run["dt"], np.array([prev]), double=False)
        assert found32[0] and abs(float(vk[0]) - s["commanded_velocity"]) < 5e-6, (k, vk[0], s["commanded_velocity"])
    assert saw_fall and saw_liftoff
    assert run["after_reset"] == 0.0

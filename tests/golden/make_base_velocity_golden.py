#!/usr/bin/env python
# SPDX-License-Identifier: Apache-2.0
"""Golden run of the reference's OWN UpkieBaseVelocity env (upkie/envs/upkie_base_velocity.py) on top of the oracle.

Run in the build container:  python tests/golden/make_base_velocity_golden.py

Combines the stand-ins of make_wrapper_golden.py (gymnasium, rate limiter, upkie_description; the physics behind the
reference's Backend ABC is oracle/) and of make_mpc_golden.py (qpmpc restated in numpy, exact BVLS instead of ProxQP).
The env class, the gyropod / servos wrappers underneath it and the MPCBalancer shell are the reference's code.
Output: tests/golden/base_velocity_run.json, replayed by tests/test_base_velocity_golden.py through
upkie_b200.base_velocity.base_velocity_tick (the function the product's env runs) with oracle-backed callables.
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "base_velocity_run.json")
sys.path.insert(0, HERE)


def main():
    import make_mpc_golden as mg
    import make_wrapper_golden as wg

    wg.install_fake_gymnasium()
    mg.install_stand_ins()
    envs = wg.load_reference()
    # the balancer module under its real name (UpkieBaseVelocity imports it lazily, upkie_base_velocity.py:117)
    import types

    ctrl = types.ModuleType("upkie.controllers")
    ctrl.__path__ = [os.path.join(wg.REF, "upkie", "controllers")]
    sys.modules["upkie.controllers"] = ctrl
    spec = importlib.util.spec_from_file_location("upkie.controllers.mpc_balancer",
                                                  os.path.join(wg.REF, "upkie/controllers/mpc_balancer.py"))
    mb = importlib.util.module_from_spec(spec)
    sys.modules["upkie.controllers.mpc_balancer"] = mb
    spec.loader.exec_module(mb)
    spec = importlib.util.spec_from_file_location("upkie.envs.upkie_base_velocity",
                                                  os.path.join(wg.REF, "upkie/envs/upkie_base_velocity.py"))
    bv = importlib.util.module_from_spec(spec)
    sys.modules["upkie.envs.upkie_base_velocity"] = bv
    spec.loader.exec_module(bv)

    import tempfile

    from oracle import oracle as O
    from upkie_b200 import _abi
    from upkie_b200.envs import spine_row_to_dict
    from upkie_b200.model import Model as B200Model
    from upkie_b200.urdf import write_urdf
    from upkie_b200.wire import action_dict_to_row

    Backend = sys.modules["upkie.envs.backends"].Backend
    RefModel = sys.modules["upkie.model"].Model
    RobotState = sys.modules["upkie.utils.robot_state"].RobotState
    b200_model = B200Model.standard_upkie()
    urdf_path = os.path.join(tempfile.mkdtemp(), "robot.urdf")
    write_urdf(b200_model, urdf_path, split_fixed_links=True)
    ref_model = RefModel(urdf_path)

    class OracleBackend(Backend):
        def __init__(self):
            self.cfg = _abi.default_sim_config()
            self.cfg.skip_action_clamps = 1
            self.sim = O.OracleSim(b200_model, self.cfg, 1, threads=1)

        def close(self):
            pass

        def get_spine_observation(self):
            return spine_row_to_dict(self.sim.spine_obs()[0])

        def reset(self, init_state):
            row = np.zeros((1, _abi.INIT_DIM))
            row[0, 0:3] = init_state.position_base_in_world
            q = init_state.orientation_base_in_world.as_quat()
            row[0, 3:7] = [q[3], q[0], q[1], q[2]]
            row[0, 7:10] = init_state.linear_velocity_base_to_world_in_world
            row[0, 10:13] = init_state.angular_velocity_base_in_base
            row[0, 13:19] = init_state.joint_configuration
            self.init_row = row.copy()
            self.sim.reset(row)
            return self.get_spine_observation()

        def step(self, action):
            a = action_dict_to_row(action).astype(np.float64)
            self.sim.step_servos(a.reshape(1, 6, 6))
            return self.get_spine_observation()

    backend = OracleBackend()
    init = RobotState(position_base_in_world=np.array([0.0, 0.0, 0.58]))
    servos = envs["upkie_servos"].UpkieServos(backend=backend, frequency=200.0, frequency_checks=False, init_state=init,
                                               regulate_frequency=False, model=ref_model)
    env = bv.UpkieBaseVelocity(servos)
    obs, info = env.reset(seed=1)
    assert not obs.any()
    rng = np.random.default_rng(20260926)
    run = {"generator": "tests/golden/make_base_velocity_golden.py", "init_row": backend.init_row[0].tolist(),
           "actions": [], "obs": [], "terminated": [], "commanded_velocity": []}
    v, w = 0.0, 0.0
    for t in range(300):  # 1.5 s: accelerate, turn, stop
        if t % 50 == 0:
            v, w = float(rng.uniform(-0.4, 0.4)), float(rng.uniform(-0.8, 0.8))
        act = np.array([v, w], dtype=np.float32)
        obs, reward, terminated, truncated, info = env.step(act)
        run["actions"].append([float(x) for x in act])
        run["obs"].append([float(x) for x in obs])
        run["terminated"].append(bool(terminated))
        run["commanded_velocity"].append(float(env.mpc_balancer.commanded_velocity))
    print("final obs", run["obs"][-1], "max |pitch| proxy: terminated any", any(run["terminated"]))
    with open(OUT, "w") as f:
        json.dump(run, f)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()

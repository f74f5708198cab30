#!/usr/bin/env python
# SPDX-License-Identifier: Apache-2.0
"""Golden trajectories of the reference's OWN env wrappers running on top of the fp64 oracle.

Run in the build container (where /root/reference exists):

    python tests/golden/make_wrapper_golden.py

The reference's ``UpkieServos`` / ``UpkieGyropod`` / ``UpkiePendulum`` classes (upkie/envs/*.py) are imported
unmodified under stand-ins for what is absent here -- a minimal ``gymnasium`` (Env, Wrapper, Box, Dict), a no-op
``loop_rate_limiters.RateLimiter`` and an empty ``upkie_description`` -- and driven through their public API
(``reset(seed)``, ``step(action)``) with ``backend = OracleBackend``: an implementation of the reference's ``Backend``
ABC (upkie/envs/backends/backend.py:11-50) whose physics is ``oracle/`` (N = 1). Everything between the agent's action
and the backend call, and between the backend's spine observation and the env observation, is therefore the
reference's code: action clamping and the spine action dictionary (upkie_servos.py:308-344), the gyropod wheel
velocity map, leg low-pass filter, yaw integration, fall detection (upkie_gyropod.py:186-392), the pendulum index map
(upkie_pendulum.py:17,124-142), reset sampling (upkie_env.py:162-194).

Outputs ``tests/golden/wrapper_trajectories.json``; ``tests/test_wrapper_golden.py`` replays the same seeds and
actions through the oracle's own restatement of those wrappers (``OracleSim.step_gyropod`` etc.) and through the
kernels' CPU build, and compares. Nothing at test time reads /root/reference.
"""

import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("UPKIE_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wrapper_trajectories.json")
sys.path.insert(0, ROOT)


# ---- stand-ins for absent third-party modules (test infrastructure only) ----------------------
def install_fake_gymnasium():
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")

    class Space:
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape if shape is not None else np.shape(low)).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape if shape is not None else np.shape(high)).copy()
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)

    class Dict(Space, dict):
        def __init__(self, spaces_=None):
            dict.__init__(self, spaces_ or {})
            self.spaces = self

    class Env:
        np_random = None

        def reset(self, seed=None, options=None):
            if seed is not None or self.np_random is None:
                self.np_random = np.random.default_rng(seed)  # gymnasium.utils.seeding.np_random -> PCG64 Generator

        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def reset(self, seed=None, options=None):
            return self.env.reset(seed=seed, options=options)

        def step(self, action):
            return self.env.step(action)

    spaces.Space, spaces.Box, spaces.Dict = Space, Box, Dict
    gym.Env, gym.Wrapper, gym.Space, gym.spaces = Env, Wrapper, Space, spaces
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = spaces
    rl = types.ModuleType("loop_rate_limiters")

    class RateLimiter:
        def __init__(self, *a, **k):
            pass

        def sleep(self):
            pass

    rl.RateLimiter = RateLimiter
    sys.modules["loop_rate_limiters"] = rl
    ud = types.ModuleType("upkie_description")
    ud.URDF_PATH = ""
    sys.modules["upkie_description"] = ud


def load_reference():
    def pkg(name, rel):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, rel)]
        sys.modules[name] = m
        return m

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    pkg("upkie", "upkie")
    pkg("upkie.utils", "upkie/utils")
    load("upkie.exceptions", "upkie/exceptions.py")
    load("upkie.logging", "upkie/logging.py")
    for m in ("filters", "clamp", "rotations", "robot_state_randomization", "robot_state", "external_force"):
        load(f"upkie.utils.{m}", f"upkie/utils/{m}.py")
    mp = pkg("upkie.model", "upkie/model")
    for m in ("se3", "joint_limit", "joint", "collision_geometry", "link", "kinematic_tree", "joint_properties", "model"):
        mod = load(f"upkie.model.{m}", f"upkie/model/{m}.py")
    mp.Model = sys.modules["upkie.model.model"].Model
    mp.JointProperties = sys.modules["upkie.model.joint_properties"].JointProperties
    pkg("upkie.envs", "upkie/envs")
    bp = pkg("upkie.envs.backends", "upkie/envs/backends")
    bp.Backend = load("upkie.envs.backends.backend", "upkie/envs/backends/backend.py").Backend
    envs = {}
    for m in ("upkie_env", "upkie_servos", "upkie_gyropod", "upkie_pendulum"):
        envs[m] = load(f"upkie.envs.{m}", f"upkie/envs/{m}.py")
    return envs


def main():
    install_fake_gymnasium()
    envs = load_reference()
    from oracle import oracle as O
    from upkie_b200 import _abi
    from upkie_b200.envs import spine_row_to_dict
    from upkie_b200.model import Model as B200Model
    from upkie_b200.urdf import write_urdf
    from upkie_b200.wire import action_dict_to_row

    Backend = sys.modules["upkie.envs.backends"].Backend
    RefModel = sys.modules["upkie.model"].Model
    RobotState = sys.modules["upkie.utils.robot_state"].RobotState
    Randomization = sys.modules["upkie.utils.robot_state_randomization"].RobotStateRandomization

    b200_model = B200Model.standard_upkie()
    tmp = tempfile.mkdtemp()
    urdf_path = os.path.join(tmp, "robot.urdf")
    write_urdf(b200_model, urdf_path, split_fixed_links=True)
    ref_model = RefModel(urdf_path)

    class OracleBackend(Backend):
        """The reference's Backend contract on top of oracle/ (one robot)."""

        def __init__(self):
            self.cfg = _abi.default_sim_config()
            self.cfg.skip_action_clamps = 1  # UpkieServos already clamped (as B200Backend does)
            self.sim = O.OracleSim(b200_model, self.cfg, 1, threads=1)
            self.spine_actions = []

        def close(self):
            pass

        def get_spine_observation(self):
            return spine_row_to_dict(self.sim.spine_obs()[0])

        def reset(self, init_state):
            row = np.zeros((1, _abi.INIT_DIM))
            row[0, 0:3] = init_state.position_base_in_world
            q = init_state.orientation_base_in_world.as_quat()  # scipy: x, y, z, w
            row[0, 3:7] = [q[3], q[0], q[1], q[2]]
            row[0, 7:10] = init_state.linear_velocity_base_to_world_in_world
            row[0, 10:13] = init_state.angular_velocity_base_in_base
            row[0, 13:19] = init_state.joint_configuration
            self.sim.reset(row)
            return self.get_spine_observation()

        def step(self, action):
            a = action_dict_to_row(action).astype(np.float64)
            self.spine_actions.append(a.tolist())
            self.sim.step_servos(a.reshape(1, 6, 6))
            return self.get_spine_observation()

    def nan_to_none(x):
        return [[None if (isinstance(v, float) and v != v) else v for v in row] for row in x]

    out = {"generator": "tests/golden/make_wrapper_golden.py",
           "reference_commit": "0a82a89b011cd179b20f572c486787ffc0ea69d6", "cases": []}
    rng = np.random.default_rng(20260924)
    UpkieServos = envs["upkie_servos"].UpkieServos
    UpkieGyropod = envs["upkie_gyropod"].UpkieGyropod
    UpkiePendulum = envs["upkie_pendulum"].UpkiePendulum
    for kind, steps, seed in (("gyropod", 60, 3), ("pendulum", 60, 4), ("gyropod", 250, 5), ("servos", 30, 6)):
        backend = OracleBackend()
        init = RobotState(
            position_base_in_world=np.array([0.0, 0.0, 0.58]),
            randomization=Randomization(pitch=0.05, x=0.1, omega_y=0.1),
        )
        servos = UpkieServos(backend=backend, frequency=200.0, frequency_checks=False, init_state=init,
                             regulate_frequency=False, model=ref_model)
        if kind == "servos":
            env = servos
        elif kind == "gyropod":
            env = UpkieGyropod(servos)
        else:
            env = UpkiePendulum(servos)
        obs, info = env.reset(seed=seed)
        case = {"kind": kind, "seed": seed, "reset_obs": None, "actions": [], "obs": [], "terminated": [],
                "init_row": None}
        st = backend.sim.get_state()[0]
        case["state_after_reset"] = st.tolist()
        if kind == "servos":
            case["reset_obs"] = [[float(obs[j][k][0]) for k in _abi.OBS_KEYS] for j in _abi.JOINT_NAMES]
        else:
            case["reset_obs"] = [float(x) for x in obs]
        for t in range(steps):
            if kind == "servos":
                act = {}
                arr = np.zeros((6, 6))
                for j, name in enumerate(_abi.JOINT_NAMES):
                    wheel = "wheel" in name
                    vals = {
                        "position": float("nan") if wheel else float(rng.uniform(-1.5, 1.5)),  # beyond the hip limit: clamped
                        "velocity": float(rng.uniform(-40.0, 40.0)) if wheel else 0.0,
                        "feedforward_torque": float(rng.uniform(-0.3, 0.3)),
                        "kp_scale": float(rng.uniform(0.0, 6.0)),  # beyond max_gain_scale = 5: clamped
                        "kd_scale": float(rng.uniform(0.0, 2.0)),
                        "maximum_torque": float(rng.uniform(0.0, 20.0)),  # beyond tau_max: clamped
                    }
                    act[name] = {k: np.array([v], dtype=np.float32) for k, v in vals.items()}
                    arr[j] = [vals[k] for k in _abi.ACT_KEYS]
                case["actions"].append(nan_to_none(np.asarray(arr, dtype=np.float32).astype(float).tolist()))
            elif kind == "gyropod":
                act = np.array([rng.uniform(-3.5, 3.5) if t % 7 else 0.0, rng.uniform(-1.2, 1.2)], dtype=np.float32)
                if seed == 5:  # a long push that makes the robot fall: termination must flip at the same step
                    act = np.array([3.0, 0.0], dtype=np.float32)
                case["actions"].append([float(x) for x in act])
            else:
                act = np.array([rng.uniform(-3.5, 3.5)], dtype=np.float32)
                case["actions"].append([float(x) for x in act])
            obs, reward, terminated, truncated, info = env.step(act)
            assert reward == 0.0 and truncated is False
            if kind == "servos":
                case["obs"].append([[float(obs[j][k][0]) for k in _abi.OBS_KEYS] for j in _abi.JOINT_NAMES])
            else:
                case["obs"].append([float(x) for x in obs])
            case["terminated"].append(bool(terminated))
        case["spine_actions"] = [nan_to_none(a) for a in backend.spine_actions]
        out["cases"].append(case)
        print(kind, "seed", seed, "steps", steps, "terminated at", [i for i, x in enumerate(case["terminated"]) if x][:3])

    # spaces and neutral action of the reference's classes (upkie_servos.py:173-286, upkie_gyropod.py:118-160,
    # upkie_pendulum.py:87-102): what B200VectorEnv's single_*_space must equal
    def box(b):
        return {"low": np.asarray(b.low, dtype=float).tolist(), "high": np.asarray(b.high, dtype=float).tolist(),
                "shape": list(b.shape), "dtype": str(np.dtype(b.dtype))}

    servos = UpkieServos(backend=OracleBackend(), frequency=200.0, frequency_checks=False, init_state=None,
                         regulate_frequency=False, model=ref_model)
    gyro = UpkieGyropod(servos)
    pend = UpkiePendulum(servos)
    out["spaces"] = {
        "servos_action": {j: {k: box(v) for k, v in servos.action_space[j].items()} for j in servos.action_space},
        "servos_observation": {j: {k: box(v) for k, v in servos.observation_space[j].items()} for j in servos.observation_space},
        "neutral_action": {j: {k: (None if (isinstance(v, float) and v != v) else float(v)) for k, v in d.items()}
                           for j, d in servos.get_neutral_action().items()},
        "gyropod_action": box(gyro.action_space), "gyropod_observation": box(gyro.observation_space),
        "pendulum_action": box(pend.action_space), "pendulum_observation": box(pend.observation_space),
    }

    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()

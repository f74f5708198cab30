#!/usr/bin/env python
# SPDX-License-Identifier: Apache-2.0
"""Golden closed-loop runs of the reference's OWN MPCBalancer shell.

Run in the build container (where /root/reference exists):

    python tests/golden/make_mpc_golden.py

``upkie/controllers/mpc_balancer.py`` is imported unmodified. Its third-party dependencies are absent here, so they
are replaced by stand-ins (TEST INFRASTRUCTURE, not the product):

* ``qpmpc`` (``WheeledInvertedPendulum``, ``MPCProblem``, ``MPCQP``, ``Plan``): a numpy restatement of the published
  formulation -- zero-order-hold discretisation of ``p'' = u, theta'' = w^2 theta - u / l``,
  condensed QP ``P = w_u I + w_T Psi_N' Psi_N + w_x sum_k Psi_k' Psi_k``, input bounds interleaved as
  ``G = [+e_k; -e_k]``, ``h = a_max`` (the interleaving the reference itself de-interleaves, mpc_balancer.py:84-86);
* ``proxsuite.proxqp.dense.QP``: exact bounded least squares (scipy BVLS on the Cholesky factor) instead of ProxQP;
* ``qpsolvers.Solution``: a record.

What IS the reference's code in these runs: the observation unpacking, ``get_target_states``, the cost-vector update
sequence, and the post-processing of ``MPCBalancer.step`` (fall / no-contact low-pass with cutoff 0.1 s,
``v += a dt / 2``, ``clamp_abs``; mpc_balancer.py:237-312). The golden commanded velocities pin the oracle's
``oracle_mpc_step`` (an independent C++ restatement) and, through it, the kernel.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("UPKIE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mpc_balancer_runs.json")


def install_stand_ins():
    from scipy.optimize import lsq_linear

    qpmpc = types.ModuleType("qpmpc")
    systems = types.ModuleType("qpmpc.systems")

    class MPCProblem:
        def __init__(self, A, B, nb_timesteps, a_max, w_T, w_x, w_u):
            self.transition_state_matrix, self.transition_input_matrix = A, B
            self.nb_timesteps, self.a_max = nb_timesteps, a_max
            self.terminal_cost_weight, self.stage_state_cost_weight, self.stage_input_cost_weight = w_T, w_x, w_u
            self.initial_state = None
            self.goal_state = None
            self.target_states = None

        def update_initial_state(self, x):
            self.initial_state = np.array(x, dtype=float)

        def update_goal_state(self, x):
            self.goal_state = np.array(x, dtype=float)

        def update_target_states(self, x):
            self.target_states = np.array(x, dtype=float)

    class WheeledInvertedPendulum:
        STATE_DIM = 4
        INPUT_DIM = 1

        def __init__(self, length, max_ground_accel, nb_timesteps, sampling_period):
            self.length, self.max_ground_accel = length, max_ground_accel
            self.nb_timesteps, self.sampling_period = nb_timesteps, sampling_period
            self.state = np.zeros(4)
            g = 9.81
            self.omega = np.sqrt(g / length)

        def build_mpc_problem(self, terminal_cost_weight, stage_state_cost_weight, stage_input_cost_weight):
            T, w, g = self.sampling_period, self.omega, 9.81
            ch, sh = np.cosh(T * w), np.sinh(T * w)
            A = np.array([[1, 0, T, 0], [0, ch, 0, sh / w], [0, 0, 1, 0], [0, w * sh, 0, ch]], dtype=float)
            B = np.array([[T * T / 2.0], [(1.0 - ch) / g], [T], [-w * sh / g]])
            return MPCProblem(A, B, self.nb_timesteps, self.max_ground_accel, terminal_cost_weight,
                              stage_state_cost_weight, stage_input_cost_weight)

    class MPCQP:
        def __init__(self, problem):
            N, A, B = problem.nb_timesteps, problem.transition_state_matrix, problem.transition_input_matrix
            self.Phi = [np.linalg.matrix_power(A, k) for k in range(N + 1)]
            self.Psi = []
            for k in range(N + 1):  # x_k = Phi_k x0 + Psi_k U
                Psi = np.zeros((4, N))
                for j in range(k):
                    Psi[:, j:j + 1] = np.linalg.matrix_power(A, k - 1 - j) @ B
                self.Psi.append(Psi)
            w_T, w_x, w_u = problem.terminal_cost_weight, problem.stage_state_cost_weight, problem.stage_input_cost_weight
            self.P = w_u * np.eye(N) + w_T * self.Psi[N].T @ self.Psi[N] + w_x * sum(self.Psi[k].T @ self.Psi[k] for k in range(N))
            self.G = np.zeros((2 * N, N))
            self.G[0::2] = np.eye(N)
            self.G[1::2] = -np.eye(N)
            self.h = np.full(2 * N, float(problem.a_max))
            self.q = np.zeros(N)
            self.problem = types.SimpleNamespace(P=self.P, q=self.q, G=self.G, h=self.h)
            self.update_cost_vector(problem)

        def update_cost_vector(self, problem):
            N = problem.nb_timesteps
            x0 = problem.initial_state
            w_T, w_x = problem.terminal_cost_weight, problem.stage_state_cost_weight
            q = w_T * self.Psi[N].T @ (self.Phi[N] @ x0 - problem.goal_state)
            for k in range(N):
                q = q + w_x * self.Psi[k].T @ (self.Phi[k] @ x0 - problem.target_states[4 * k:4 * k + 4])
            self.q = q
            self.problem.q = q

    class Plan:
        def __init__(self, problem, qpsol):
            self.inputs = None if (qpsol is None or qpsol.x is None) else np.asarray(qpsol.x).reshape(-1, 1)

        @property
        def is_empty(self):
            return self.inputs is None

        @property
        def first_input(self):
            return self.inputs[0]

    qpmpc.MPCQP, qpmpc.MPCProblem, qpmpc.Plan = MPCQP, MPCProblem, Plan
    systems.WheeledInvertedPendulum = WheeledInvertedPendulum
    qpmpc.systems = systems
    sys.modules["qpmpc"] = qpmpc
    sys.modules["qpmpc.systems"] = systems

    proxsuite = types.ModuleType("proxsuite")
    proxqp = types.SimpleNamespace()
    dense = types.SimpleNamespace()

    class QP:
        def __init__(self, n, n_eq, n_in, dense_backend=None):
            self.settings = types.SimpleNamespace()
            self.results = types.SimpleNamespace(x=None, info=types.SimpleNamespace(status=None))

        def init(self, H, g, C, l, u):  # noqa: E741
            self.H, self.g, self.l, self.u = np.array(H), np.array(g), np.array(l), np.array(u)
            assert np.array_equal(C, np.eye(len(g)))  # box on the inputs
            self.L = np.linalg.cholesky(self.H)

        def update(self, g=None, update_preconditioner=True):
            self.g = np.array(g)

        def solve(self):
            r = lsq_linear(self.L.T, -np.linalg.solve(self.L, self.g), bounds=(self.l, self.u), method="bvls", tol=1e-14)
            self.results.x = r.x
            self.results.info.status = "solved"

    dense.QP = QP
    dense.DenseBackend = types.SimpleNamespace(PrimalDualLDLT=0)
    proxqp.dense = dense
    proxqp.QPSolverOutput = types.SimpleNamespace(PROXQP_SOLVED="solved")
    proxsuite.proxqp = proxqp
    sys.modules["proxsuite"] = proxsuite
    qpsolvers = types.ModuleType("qpsolvers")

    class Solution:
        def __init__(self, problem):
            self.problem, self.found, self.x = problem, False, None

    qpsolvers.Solution = Solution
    sys.modules["qpsolvers"] = qpsolvers


def load_reference_balancer():
    def pkg(name, rel):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, rel)]
        sys.modules[name] = m

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
    load("upkie.utils.filters", "upkie/utils/filters.py")
    load("upkie.utils.clamp", "upkie/utils/clamp.py")
    return load("upkie_mpc_balancer", "upkie/controllers/mpc_balancer.py")


def main():
    install_stand_ins()
    mod = load_reference_balancer()
    rng = np.random.default_rng(20260925)
    out = {"generator": "tests/golden/make_mpc_golden.py", "reference_commit": "0a82a89b011cd179b20f572c486787ffc0ea69d6", "runs": []}
    for horizon in (16, 50):
        balancer = mod.MPCBalancer(nb_timesteps=horizon)
        run = {"nb_timesteps": horizon, "dt": 0.005, "steps": []}
        p = 0.0
        for k in range(120):
            pitch = float(rng.uniform(-0.25, 0.25))
            if 40 <= k < 46:
                pitch = float(rng.uniform(1.05, 1.3))  # fallen: low-pass towards zero
            contact = not (70 <= k < 78)  # lifted off for a few cycles
            gv = float(rng.uniform(-0.6, 0.6))
            p += gv * 0.005
            obs = {
                "floor_contact": {"contact": contact},
                "base_orientation": {"pitch": pitch, "angular_velocity": [0.0, float(rng.uniform(-1.0, 1.0)), 0.0]},
                "wheel_odometry": {"position": p, "velocity": gv},
            }
            target = float(rng.uniform(-1.0, 1.0)) if k % 10 else 0.0
            v = balancer.step(target, obs, 0.005)
            run["steps"].append({"x0": [p, pitch, gv, obs["base_orientation"]["angular_velocity"][1]], "target": target,
                                 "contact": contact, "commanded_velocity": float(v),
                                 "first_input": float(balancer.proxqp.solver.results.x[0])})
        balancer.reset()
        run["after_reset"] = float(balancer.commanded_velocity)
        out["runs"].append(run)
        print("horizon", horizon, "final commanded velocity", run["steps"][-1]["commanded_velocity"])
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()

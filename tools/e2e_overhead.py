# SPDX-License-Identifier: Apache-2.0
"""Developer tool: where the host time of B200VectorEnv.step goes on bench.py's servos workload
(library call vs Python around it). Run on the GPU box: python tools/e2e_overhead.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from upkie_b200.envs import B200VectorEnv  # noqa: E402
from upkie_b200.model import Model  # noqa: E402


def main():
    n = 65536
    dev = torch.device("cuda", 0)
    model = Model.standard_upkie()
    cfg = bench.servos_config()
    env = B200VectorEnv(n, "servos", config=cfg, device=0, autoreset_mode="next_step", model=model)
    gen = torch.Generator(device=dev)
    gen.manual_seed(2025)
    env.sim.set_randomization(friction=torch.empty(n, device=dev).uniform_(0.5, 1.2, generator=gen),
                              inertia_eps=torch.empty((n, 6), device=dev).uniform_(-0.2, 0.2, generator=gen))
    env.sim.set_autoreset(1, 2025, 0)
    env.sim.reset(seed=2025)
    tau = torch.tensor(model.tau_max, dtype=torch.float32, device=dev)
    acts = []
    for _ in range(4):
        a = torch.zeros((n, 6, 6), device=dev)
        a[:, :, 0] = float("nan")
        a[:, :, 5] = tau
        a[:, :, 2] = (torch.rand((n, 6), device=dev, generator=gen) * 2 - 1) * tau
        acts.append(a.cpu().pin_memory().numpy())
    for name, fn in (("sim.step_servos_host_compact", env.sim.step_servos_host_compact), ("env.step", env.step)):
        for k in range(20):
            fn(acts[k % 4])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(300):
            fn(acts[k % 4])
        dt = (time.perf_counter() - t0) / 300
        print(f"{name}: {dt * 1e3:.4f} ms/step  {n / dt / 1e6:.1f} M env-steps/s", flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for k in range(300):
        env.step(acts[k % 4])
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)


if __name__ == "__main__":
    main()

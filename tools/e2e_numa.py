# SPDX-License-Identifier: Apache-2.0
"""Developer tool: effect of the host thread / pinned-buffer NUMA placement on the host-buffer step.
Run on the GPU box: python tools/e2e_numa.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upkie_b200 import _abi, numa  # noqa: E402
from upkie_b200.model import Model  # noqa: E402
from upkie_b200.sim import UpkieSim  # noqa: E402


def rate(n, m, label):
    sim = UpkieSim(n, model=m, config=_abi.default_sim_config())
    sim.reset(seed=1)
    act = sim.host_action_buffer(36)
    act[:] = 0
    act[:, :, 0] = float("nan")
    act[:, :, 3:5] = 1.0
    act[:, :, 5] = m.tau_max
    for _ in range(30):
        sim.step_servos_host_compact(act)
    torch.cuda.synchronize()
    best = 1e9
    tot = 0.0
    for rep in range(5):
        t0 = time.perf_counter()
        for _ in range(100):
            sim.step_servos_host_compact(act)
        dt = (time.perf_counter() - t0) / 100
        best = min(best, dt)
        tot += dt
    print(f"{label}: mean {tot / 5 * 1e3:.4f} ms  best {best * 1e3:.4f} ms  ({n / (tot / 5) / 1e6:.1f} M env-steps/s)", flush=True)
    sim.close()


def main():
    n = 65536
    m = Model.standard_upkie()
    torch.cuda.init()
    node = numa.gpu_numa_node(0)
    all_cpus = sorted(os.sched_getaffinity(0))
    print("gpu numa node", node, "cpus", len(all_cpus), "nodes", sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")))
    rate(n, m, "unbound")
    if node is not None:
        for nd in sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node")):
            with open(f"/sys/devices/system/node/node{nd}/cpulist") as f:
                cpus = sorted(set(numa._parse_cpulist(f.read())) & set(all_cpus))
            if not cpus:
                continue
            os.sched_setaffinity(0, cpus)
            rate(n, m, f"bound to node {nd}" + (" (GPU's node)" if nd == node else ""))
        os.sched_setaffinity(0, all_cpus)
    rate(n, m, "unbound again")


if __name__ == "__main__":
    main()

# SPDX-License-Identifier: Apache-2.0
"""Where the end-to-end step time goes: pinned H2D / D2H bandwidth at the step's sizes (alone and
full-duplex), and the step kernel's duration as a function of the number of envs per launch.
Developer tool (run on the GPU box): python tools/e2e_parts.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upkie_b200 import _abi  # noqa: E402
from upkie_b200.model import Model  # noqa: E402
from upkie_b200.sim import UpkieSim  # noqa: E402


def timed(fn, reps=50):
    torch.cuda.synchronize()
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = 65536
    m = Model.standard_upkie()
    # kernel time vs envs per launch (standing robots, PD action)
    for cnt in ((4096, 8192, 16384, 32768, 65536) if "--kernel" in sys.argv else ()):
        sim = UpkieSim(cnt, model=m, config=_abi.default_sim_config())
        sim.reset(seed=1)
        a = torch.zeros((cnt, 6, 6), device="cuda")
        a[:, :, 0] = float("nan")
        a[:, :, 3:5] = 1.0
        a[:, :, 5] = torch.tensor(m.tau_max, device="cuda", dtype=torch.float32)
        for _ in range(30):
            sim.step_servos(a)
        t = timed(lambda: sim.step_servos(a), reps=100)
        print(f"kernel {cnt} envs: {t:.4f} ms  {cnt / t / 1e3:.1f} M env-steps/s")
    # host path as shipped: zero-copy persistent kernel (block, blocks/SM) and the staged pipeline (chunks)
    import time
    configs = []
    for split, k in (("", 1), ("0.4,0.4,0.2", 2), ("0.45,0.35,0.2", 2), ("0.6,0.4", 1), ("0.6,0.4", 2), ("0.4,0.35,0.25", 2),
                     ("0.35,0.3,0.2,0.15", 2), ("0.5,0.3,0.2", 2), ("0.5,0.3,0.2", 1), ("0.3,0.3,0.25,0.15", 2)):
        c = {"compact": "1", "UPKIE_B200_ZERO_COPY": "2", "UPKIE_B200_HOST_KERNEL_STREAMS": str(k)}
        if split:
            c["UPKIE_B200_HOST_SPLIT"] = split
        configs.append(c)
    for cfg in (configs if "--host" in sys.argv or "--kernel" not in sys.argv else []):
        for k in ("UPKIE_B200_ZERO_COPY", "UPKIE_B200_HOST_BLOCK", "UPKIE_B200_HOST_BLOCKS_PER_SM", "UPKIE_B200_HOST_CHUNKS",
                  "UPKIE_B200_HOST_KERNEL_STREAMS", "UPKIE_B200_HOST_SPLIT"):
            os.environ.pop(k, None)
        os.environ.update({k: v for k, v in cfg.items() if k != "compact"})
        sim = UpkieSim(n, model=m, config=_abi.default_sim_config())
        sim.reset(seed=1)
        act = sim.host_action_buffer(36)
        act[:] = 0
        act.reshape(n, 6, 6)[:, :, 0] = np.nan
        act.reshape(n, 6, 6)[:, :, 3:5] = 1.0
        act.reshape(n, 6, 6)[:, :, 5] = m.tau_max
        step = sim.step_servos_host_compact if cfg.get("compact") else sim.step_servos_host
        for _ in range(30):
            step(act)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            step(act)
        dt = (time.perf_counter() - t0) / 200
        print(f"step_servos_host {cfg}: {dt * 1e3:.4f} ms  {n / dt / 1e6:.1f} M env-steps/s", flush=True)


if __name__ == "__main__":
    main()

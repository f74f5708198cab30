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
    for nbytes, name in ((n * 144, "H2D action 144 B/env"), (n * 126, "D2H obs+flags 126 B/env"), (n * 73, "D2H compact 73 B/env")):
        h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        if name.startswith("H2D"):
            t = timed(lambda: d.copy_(h, non_blocking=True))
        else:
            t = timed(lambda: h.copy_(d, non_blocking=True))
        print(f"{name}: {nbytes / 1e6:.2f} MB  {t:.4f} ms  {nbytes / t / 1e6:.1f} GB/s")
    # full duplex
    ha = torch.empty(n * 144, dtype=torch.uint8).pin_memory()
    da = torch.empty(n * 144, dtype=torch.uint8, device="cuda")
    ho = torch.empty(n * 126, dtype=torch.uint8).pin_memory()
    do = torch.empty(n * 126, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def duplex():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            da.copy_(ha, non_blocking=True)
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    print(f"duplex H2D 9.4 MB + D2H 8.3 MB: {timed(duplex):.4f} ms")
    # chunked H2D: 4 copies of a quarter
    def h2d_chunks(k):
        q = n * 144 // k
        for c in range(k):
            da[c * q:(c + 1) * q].copy_(ha[c * q:(c + 1) * q], non_blocking=True)
    for k in (1, 2, 4, 8):
        print(f"H2D in {k} chunks: {timed(lambda: h2d_chunks(k)):.4f} ms")

    # kernel time vs envs per launch (standing robots, PD action)
    for cnt in (4096, 8192, 16384, 32768, 65536):
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
    # host path as shipped
    for chunks in (1, 2, 3, 4, 6, 8):
        os.environ["UPKIE_B200_HOST_CHUNKS"] = str(chunks)
        sim = UpkieSim(n, model=m, config=_abi.default_sim_config())
        sim.reset(seed=1)
        act = sim.host_action_buffer(36)
        act[:] = 0
        act.reshape(n, 6, 6)[:, :, 0] = np.nan
        act.reshape(n, 6, 6)[:, :, 3:5] = 1.0
        act.reshape(n, 6, 6)[:, :, 5] = m.tau_max
        for _ in range(30):
            sim.step_servos_host(act)
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            sim.step_servos_host(act)
        dt = (time.perf_counter() - t0) / 200
        print(f"step_servos_host, {chunks} chunks: {dt * 1e3:.4f} ms  {n / dt / 1e6:.1f} M env-steps/s")


if __name__ == "__main__":
    main()

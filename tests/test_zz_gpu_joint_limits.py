# SPDX-License-Identifier: Apache-2.0
"""Joint-limit rows on the device (k_step<.., NOISE=2, ..>, config.joint_limits = 1) against the oracle.

This file sorts last on purpose: the "extras + limits" kernels were written after round 1's GPU budget had been
spent, so their first run on a B200 is the driver's round-end `pytest -m gpu`. The CPU build of
the same code agrees with the oracle (tests/test_kernel_arithmetic_cpu.py::test_joint_limit_rows)."""
import numpy as np
import pytest

from conftest import at_joint_bounds, random_servo_actions
from upkie_b200 import _abi

pytestmark = pytest.mark.gpu


def test_joint_limit_rows_on_device(model, oracle_lib):
    import torch

    from upkie_b200.sim import UpkieSim

    n = 2048
    cfg = _abi.default_sim_config()
    cfg.joint_limits = 1
    sim = UpkieSim(n, model=model, config=cfg)
    plain = UpkieSim(n, model=model, config=_abi.default_sim_config())
    osim = oracle_lib.OracleSim(model, cfg, n, threads=8)
    st = at_joint_bounds(model, n, seed=5)
    act = random_servo_actions(n, model, seed=12).astype(np.float32)
    a = torch.from_numpy(act).cuda()
    sim.set_state(torch.from_numpy(st).cuda())
    plain.set_state(torch.from_numpy(st).cuda())
    osim.set_state(st.astype(np.float64))
    for _ in range(3):
        sim.step_servos(a)
        plain.step_servos(a)
        osim.step_servos(act.astype(np.float64))
        g = sim.get_state().cpu().numpy().astype(np.float64)
        o = osim.get_state()
        d = np.abs(g[:, :25] - o[:, :25])
        assert d[:, :7].max() < 2e-5 and d[:, 13:19].max() < 2e-4
        assert np.median(d[:, 19:25].max(axis=1)) < 5e-5 and np.percentile(d[:, 19:25].max(axis=1), 99) < 2e-3
        assert np.array_equal(g[:, 40], o[:, 40])  # contact flags
        assert np.abs(g[:, 19:25] - plain.get_state().cpu().numpy()[:, 19:25]).max() > 5.0  # the rows matter
        osim.set_state(g)
        plain.set_state(sim.get_state())

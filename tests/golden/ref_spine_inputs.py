# SPDX-License-Identifier: Apache-2.0
"""Seeded input sequences shared by tests/golden/make_ref_spine_golden.py (which feeds them to the reference's own
C++ observers / controllers) and tests/test_ref_spine.py (which feeds them to the oracle and the kernel arithmetic).
numpy's PCG64 streams are reproducible across machines; only the OUTPUTS are stored in ref_spine_runs.json."""
import numpy as np

SEED = 20260928
N_STEPS = 240


def _quat_wxyz_from_zyx(yaw, pitch, roll):
    cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
    return [cy * cp * cr + sy * sp * sr, cy * cp * sr - sy * sp * cr, cy * sp * cr + sy * cp * sr, sy * cp * cr - cy * sp * sr]


def observer_inputs(A, stream: int):
    """``[N_STEPS, SPINE_DIM]`` spine rows: drifting IMU attitude, wheel velocity / torque bursts (touchdown) and quiet
    phases (lift-off), leg torques around the floor-contact threshold."""
    rng = np.random.default_rng([SEED, stream])
    rows = np.zeros((N_STEPS, A.SPINE_DIM))
    for k in range(N_STEPS):
        r = rows[k]
        pitch = 0.3 * np.sin(0.05 * k) + 0.02 * rng.normal()
        r[A.SP_IMU_QUAT:A.SP_IMU_QUAT + 4] = _quat_wxyz_from_zyx(0.01 * k, pitch, 0.05 * np.sin(0.02 * k))
        r[A.SP_IMU_ANGVEL:A.SP_IMU_ANGVEL + 3] = rng.normal(0.0, 0.5, 3)
        busy = (k // 60) % 2 == 1
        for j in range(6):
            so = r[A.SP_SERVO + 5 * j:A.SP_SERVO + 5 * j + 5]
            wheel = j in (2, 5)
            so[0] = rng.normal()
            so[1] = (rng.normal(0.0, 30.0) if busy else rng.normal(0.0, 0.05)) if wheel else rng.normal(0.0, 0.5)
            so[2] = (rng.normal(0.0, 0.8) if busy else rng.normal(0.0, 0.002)) if wheel else rng.normal(0.0, 6.0 if busy else 1.0)
            so[3], so[4] = 42.0, 18.0
    return rows


def controller_inputs(stream: int):
    """List of ``(obs3, target2 or None, action[6, 6])``: pitch with a fall phase, a lift-off phase, random targets."""
    rng = np.random.default_rng([SEED, 100 + stream])
    out = []
    pos = 0.0
    for k in range(N_STEPS):
        pitch = float(0.2 * np.sin(0.03 * k) + 0.02 * rng.normal())
        if 100 <= k < 110:
            pitch = 1.2
        contact = 0.0 if 160 <= k < 185 else 1.0
        pos += float(rng.normal(0.0, 0.002))
        t = rng.uniform(-1.0, 1.0, 2) * np.array([1.0, 0.3])
        target = t if k % 3 else None
        act = np.zeros((6, 6))
        act[:, 0] = rng.uniform(-0.5, 0.5, 6)
        act[:, 1] = rng.uniform(-1.0, 1.0, 6)
        act[:, 2] = rng.uniform(-0.2, 0.2, 6)
        act[:, 3:5] = 1.0
        act[:, 5] = [16.0, 16.0, 1.7, 16.0, 16.0, 1.7]
        out.append((np.array([pitch, contact, pos]), target, act))
    return out

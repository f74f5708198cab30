# SPDX-License-Identifier: Apache-2.0
"""msgpack wire format of the agent <-> spine mailbox (upkie_b200/wire.py).

Pinned on bytes produced in this container by the reference's own `serialize` + msgpack.Packer settings
(tests/golden/make_golden.py writes them into tests/golden/reference_vectors.json); the round trips then close
the loop dictionary -> bytes -> dictionary -> flat row."""
import json
import os

import numpy as np
import pytest

from upkie_b200 import _abi, wire
from upkie_b200.envs import spine_row_to_dict

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")


def _example_action():
    a = np.zeros((6, 6))
    a[:, 0] = [0.1, -0.2, np.nan, 0.3, -0.4, np.nan]
    a[:, 1] = [0.0, 0.5, -7.0, 0.25, 0.0, 7.0]
    a[:, 2] = [0.0, 0.0, 0.125, 0.0, 0.0, -0.125]
    a[:, 3] = 1.0
    a[:, 4] = [1.0, 1.0, 0.5, 1.0, 1.0, 0.5]
    a[:, 5] = [16.0, 16.0, 1.7, 16.0, 16.0, 1.7]
    return a


def test_bytes_match_the_reference_packer():
    g = json.load(open(GOLDEN))["wire"]
    d = wire.action_row_to_dict(_example_action())
    assert wire.pack_dict(d).hex() == g["action_hex"]
    nested = {"imu": {"linear_acceleration": np.array([0.5, -1.5, 9.81])}, "n": 3, "flag": True, "name": "upkie"}
    assert wire.pack_dict(nested).hex() == g["nested_hex"]
    assert wire.unpack_dict(bytes.fromhex(g["nested_hex"])) == {
        "imu": {"linear_acceleration": [0.5, -1.5, 9.81]}, "n": 3, "flag": True, "name": "upkie"}


def test_round_trips_and_mailbox_frame():
    a = _example_action()
    d = wire.unpack_dict(wire.pack_dict(wire.action_row_to_dict(a)))
    row = wire.action_dict_to_row(d)
    assert np.array_equal(np.isnan(row), np.isnan(a)) and np.allclose(np.nan_to_num(row), np.nan_to_num(a), atol=1e-6)
    # backend defaults and absent joints (pybullet_backend.py:280-291)
    partial = {"servo": {"left_wheel": {"position": float("nan"), "velocity": 2.0, "maximum_torque": 1.0}, "tail": {}}}
    r = wire.action_dict_to_row(partial)
    assert r[2, 1] == 2.0 and r[2, 2] == 0.0 and r[2, 3] == 1.0 and r[2, 4] == 1.0 and r[2, 5] == 1.0
    assert (r[[0, 1, 3, 4, 5], 3:] == 0).all()
    assert np.isnan(wire.action_dict_to_row({})[:, 0]).all()
    # observation: row -> dict -> bytes -> dict -> row
    rng = np.random.default_rng(0)
    spine = rng.normal(size=_abi.SPINE_DIM).astype(np.float32)
    spine[_abi.SP_CONTACT] = 1.0
    back = wire.observation_dict_to_row(wire.unpack_dict(wire.pack_observation(spine)))
    assert np.allclose(back, spine, rtol=0, atol=0)
    assert wire.unpack_dict(wire.pack_observation(spine)) == json.loads(json.dumps(
        spine_row_to_dict(spine), default=lambda o: o.tolist()))
    # mailbox frame [request][size][payload]
    payload = wire.pack_dict({"a": 1})
    req, body = wire.parse_frame(wire.frame(wire.Request.kAction, payload) + b"\0" * 16)
    assert req == wire.Request.kAction and body == payload
    assert wire.parse_frame(wire.frame(wire.Request.kStop)) == (wire.Request.kStop, b"")
    with pytest.raises(ValueError):
        wire.unpack_dict(payload + payload)

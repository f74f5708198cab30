# SPDX-License-Identifier: Apache-2.0
"""Multi-process path on CPU: env-index sharding + rollout all-gather over gloo
(world_size 2), the same code the GPU ranks run over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from upkie_b200.sharding import RolloutBuffer, shard_range


def test_shard_range_partitions_every_env_once():
    for n, g in [(524288, 8), (65536, 1), (10, 3), (7, 8), (4096, 2)]:
        seen = []
        for r in range(g):
            off, cnt = shard_range(n, r, g)
            seen.extend(range(off, off + cnt))
        assert seen == list(range(n))
    assert shard_range(524288, 3, 8) == (3 * 65536, 65536)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, T, obs_dim, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, cnt = shard_range(n_global, rank, world)
        buf = RolloutBuffer(T, cnt, obs_dim, "cpu")
        idx = torch.arange(off, off + cnt, dtype=torch.float32)
        for t in range(T):
            # every value is a function of (t, global env index) only; the producer writes INTO the slot
            o, r, te, tr = buf.slot(t)
            o.copy_(idx[:, None] * 10.0 + torch.arange(obs_dim)[None, :] + 1000.0 * t)
            r.copy_(idx * 0.5 + t)
            te.copy_(((idx.long() + t) % 7 == 0).to(torch.uint8))
            tr.copy_(((idx.long() + t) % 11 == 0).to(torch.uint8))
        full = buf.gather()
        assert full["obs"].shape == (T, n_global, obs_dim) and full["terminated"].dtype == torch.uint8
        np.save(os.path.join(out_dir, f"obs_{rank}.npy"), full["obs"].numpy())
        gidx = torch.arange(n_global, dtype=torch.float32)
        for t in range(T):
            assert torch.equal(full["obs"][t], gidx[:, None] * 10.0 + torch.arange(obs_dim)[None, :] + 1000.0 * t)
            assert torch.equal(full["reward"][t], gidx * 0.5 + t)
            assert torch.equal(full["terminated"][t], ((gidx.long() + t) % 7 == 0).to(torch.uint8))
            assert torch.equal(full["truncated"][t], ((gidx.long() + t) % 11 == 0).to(torch.uint8))
        raw = buf.gather_raw()
        assert raw.shape == (world, buf.nbytes) and torch.equal(raw[rank], buf.raw)
        # compact records (observation rows + terminated only: 73 B/env/step for UpkieServos' 18 floats)
        cb = RolloutBuffer(T, cnt, obs_dim, "cpu", compact=True)
        assert cb.rec == 4 * obs_dim + 1 and cb.nbytes < buf.nbytes
        for t in range(T):
            o, r, te, tr = cb.slot(t)
            assert r is None and tr is None
            o.copy_(idx[:, None] * 10.0 + torch.arange(obs_dim)[None, :] + 1000.0 * t)
            te.copy_(((idx.long() + t) % 7 == 0).to(torch.uint8))
        cfull = cb.gather()
        assert set(cfull) == {"obs", "terminated"}
        assert torch.equal(cfull["obs"], full["obs"]) and torch.equal(cfull["terminated"], full["terminated"])
        # scalar statistics reduce across ranks (mask counts)
        c = buf.terminated.sum().to(torch.float64).reshape(1)
        dist.all_reduce(c)
        assert c.item() == full["terminated"].sum().item()
    finally:
        dist.destroy_process_group()


def test_rollout_all_gather_world_size_2(tmp_path):
    world, n_global, T, obs_dim = 2, 64, 4, 30
    mp.spawn(_worker, args=(world, _free_port(), n_global, T, obs_dim, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "obs_0.npy")
    b = np.load(tmp_path / "obs_1.npy")
    assert np.array_equal(a, b)  # every rank holds the same global rollout


def test_rollout_buffer_single_process_roundtrip():
    T, n, d = 3, 5, 4
    buf = RolloutBuffer(T, n, d, "cpu")
    rng = np.random.default_rng(0)
    ref = []
    for t in range(T):
        obs = torch.from_numpy(rng.normal(size=(n, d)).astype(np.float32))
        rew = torch.zeros(n)
        term = torch.from_numpy((rng.uniform(size=n) < 0.3).astype(np.uint8))
        trunc = torch.zeros(n, dtype=torch.uint8)
        buf.record(t, obs, rew, term, trunc)
        ref.append((obs, term))
    full = buf.gather()
    for t in range(T):
        assert torch.equal(full["obs"][t], ref[t][0]) and torch.equal(full["terminated"][t], ref[t][1])
    assert buf.rec == 4 * d + 6 and RolloutBuffer(1, 1, 30, "cpu").rec == 126  # 126 B/env/step
    assert RolloutBuffer(1, 1, 18, "cpu", compact=True).rec == 73
    assert buf.nbytes >= T * n * buf.rec
    # slots are views of the one byte buffer that gets gathered
    o, r, te, tr = buf.slot(1)
    o.fill_(7.0)
    assert (buf.gather()["obs"][1] == 7.0).all()

# SPDX-License-Identifier: Apache-2.0
"""Env-index sharding across GPUs and the rollout all-gather.

Every env is independent, so the path shards by env index with no collective
inside ``step()``: rank ``r`` of ``G`` owns the contiguous block
``shard_range(N, r, G)`` and passes its ``offset`` as ``env_offset`` so that the
device RNG is keyed on the *global* env index (results do not depend on G).
The only exchange is one all-gather of the trajectory buffer per rollout, over
NCCL (NVLink/NVSwitch) on GPUs or gloo in the CPU tests.
"""

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """``(offset, count)`` of the contiguous env block owned by ``rank``.
    The first ``num_envs % world_size`` ranks take one extra env."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, rem = divmod(int(num_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


class RolloutBuffer:
    """One rollout of this rank: ``obs[T, n, obs_dim]`` f32, ``reward[T, n]`` f32,
    ``terminated[T, n]`` u8, ``truncated[T, n]`` u8, carved out of ONE byte buffer
    (``4 * obs_dim + 6`` bytes per env and step: 126 B for UpkieServos).

    ``slot(t)`` returns the four views of time step ``t``; handing them to
    ``UpkieSim.step_*`` as output tensors makes the step kernel write the rollout
    in place (no copy kernels). ``gather()`` all-gathers the byte buffer once and
    returns the rollout of ALL envs ordered by global env index, identical on
    every rank.
    """

    def __init__(self, horizon: int, n_local: int, obs_dim: int, device, group: Optional[dist.ProcessGroup] = None):
        self.T, self.n, self.obs_dim = int(horizon), int(n_local), int(obs_dim)
        self.rec = 4 * self.obs_dim + 6
        self.group = group
        self.device = torch.device(device)
        T, n, d = self.T, self.n, self.obs_dim
        self._sizes = [T * n * d * 4, T * n * 4, T * n, T * n]
        total = sum(self._sizes)
        self.nbytes = (total + 15) // 16 * 16
        self.raw = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
        self.obs, self.reward, self.terminated, self.truncated = self._views(self.raw, n)

    def _views(self, raw: torch.Tensor, n: int):
        T, d = self.T, self.obs_dim
        o0 = 0
        o1 = o0 + T * n * d * 4
        o2 = o1 + T * n * 4
        o3 = o2 + T * n
        obs = raw[o0:o1].view(torch.float32).view(T, n, d)
        rew = raw[o1:o2].view(torch.float32).view(T, n)
        term = raw[o2:o3].view(T, n)
        trunc = raw[o3:o3 + T * n].view(T, n)
        return obs, rew, term, trunc

    def slot(self, t: int):
        """Output tensors of time step ``t`` (modulo the horizon)."""
        k = t % self.T
        return self.obs[k], self.reward[k], self.terminated[k], self.truncated[k]

    def record(self, t: int, obs: torch.Tensor, reward: torch.Tensor, terminated: torch.Tensor, truncated: torch.Tensor):
        """Copy-in variant of ``slot`` for producers that own their output tensors."""
        o, r, te, tr = self.slot(t)
        o.copy_(obs.reshape(self.n, self.obs_dim))
        r.copy_(reward)
        te.copy_(terminated)
        tr.copy_(truncated)

    def gather(self) -> Dict[str, torch.Tensor]:
        """``{"obs": [T, N, obs_dim], "reward": [T, N], "terminated": [T, N], "truncated": [T, N]}``
        over all ranks (``N = world_size * n``; equal shards)."""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return {"obs": self.obs, "reward": self.reward, "terminated": self.terminated, "truncated": self.truncated}
        world = dist.get_world_size(self.group)
        out = torch.empty(world * self.nbytes, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(out, self.raw, group=self.group)
        parts = [self._views(out[r * self.nbytes:(r + 1) * self.nbytes], self.n) for r in range(world)]
        return {
            "obs": torch.cat([p[0] for p in parts], dim=1),
            "reward": torch.cat([p[1] for p in parts], dim=1),
            "terminated": torch.cat([p[2] for p in parts], dim=1),
            "truncated": torch.cat([p[3] for p in parts], dim=1),
        }

    def gather_raw(self, async_op: bool = False):
        """The all-gather alone (``[world, nbytes]`` bytes, rank-major), without re-assembly.

        With ``async_op=True`` returns ``(tensor, work)``: the collective runs on the
        communication stream while the next rollout (in another buffer) is simulated;
        call ``work.wait()`` before touching ``tensor`` or re-using this buffer's slots."""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            out = self.raw.view(1, -1)
            return (out, None) if async_op else out
        world = dist.get_world_size(self.group)
        if getattr(self, "_gathered", None) is None:
            self._gathered = torch.empty(world * self.nbytes, dtype=torch.uint8, device=self.device)
        work = dist.all_gather_into_tensor(self._gathered, self.raw, group=self.group, async_op=async_op)
        out = self._gathered.view(world, self.nbytes)
        return (out, work) if async_op else out

# SPDX-License-Identifier: Apache-2.0
"""Env-index sharding across GPUs and the rollout all-gather.

Every env is independent, so the path shards by env index with no collective
inside ``step()``: rank ``r`` of ``G`` owns the contiguous block
``shard_range(N, r, G)`` and passes its ``offset`` as ``env_offset`` so that the
device RNG is keyed on the *global* env index (results do not depend on G).
The only exchange is one all-gather of the trajectory buffer per rollout, over
NCCL (NVLink/NVSwitch) on GPUs or gloo in the CPU tests.
"""

from typing import Optional, Tuple

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
    """``[T, n_local, record]`` byte buffer of one rollout on this rank.

    A record is ``obs (obs_dim f32) | reward (f32) | terminated (u8) | truncated
    (u8)`` = ``4 * obs_dim + 6`` bytes (126 B for UpkieServos). ``gather()``
    returns the rollout of ALL envs, ``[T, N_global, record]``, ordered by global
    env index, identical on every rank.
    """

    def __init__(self, horizon: int, n_local: int, obs_dim: int, device, group: Optional[dist.ProcessGroup] = None):
        self.T, self.n, self.obs_dim = int(horizon), int(n_local), int(obs_dim)
        self.rec = 4 * self.obs_dim + 6
        self.group = group
        self.device = torch.device(device)
        self.data = torch.zeros((self.T, self.n, self.rec), dtype=torch.uint8, device=self.device)

    def record(self, t: int, obs: torch.Tensor, reward: torch.Tensor, terminated: torch.Tensor, truncated: torch.Tensor):
        r = self.data[t % self.T]
        ob = 4 * self.obs_dim
        r[:, :ob] = obs.reshape(self.n, self.obs_dim).contiguous().view(torch.uint8).reshape(self.n, ob)
        r[:, ob:ob + 4] = reward.contiguous().view(torch.uint8).reshape(self.n, 4)
        r[:, ob + 4] = terminated
        r[:, ob + 5] = truncated

    def gather(self) -> torch.Tensor:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return self.data
        world = dist.get_world_size(self.group)
        # concatenated layout [G * T, n, rec] (the form both NCCL and gloo accept)
        out = torch.empty((world * self.T, self.n, self.rec), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(out, self.data, group=self.group)
        # [G, T, n, rec] -> [T, G * n, rec]: rank-major = global env order for equal shards
        return out.reshape(world, self.T, self.n, self.rec).permute(1, 0, 2, 3).reshape(self.T, world * self.n, self.rec)

    @staticmethod
    def unpack(records: torch.Tensor, obs_dim: int):
        """Split gathered records back into ``obs[T, N, obs_dim]`` (f32),
        ``reward[T, N]`` (f32), ``terminated[T, N]``, ``truncated[T, N]`` (u8)."""
        T, N, rec = records.shape
        ob = 4 * obs_dim
        obs = records[:, :, :ob].contiguous().view(torch.float32).reshape(T, N, obs_dim)
        rew = records[:, :, ob:ob + 4].contiguous().view(torch.float32).reshape(T, N)
        return obs, rew, records[:, :, ob + 4], records[:, :, ob + 5]

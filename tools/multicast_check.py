# SPDX-License-Identifier: Apache-2.0
"""Validation of the in-kernel rollout transports (NVSwitch multicast stores, peer stores) on 2..8 GPUs.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/multicast_check.py

Every rank steps two identical simulators for T ticks: one writes compact rows into a plain local buffer
(upkie_b200_step_servos_compact), the other stores them to the multicast address of its slot in a symmetric buffer
(upkie_b200_step_servos_multicast). After publish() (cross-rank barrier) every rank must hold, for EVERY rank r, slot r
bit-identical to what rank r computed locally (exchanged with a plain all_gather for the comparison)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upkie_b200 import _abi  # noqa: E402
from upkie_b200.model import Model  # noqa: E402
from upkie_b200.sharding import PeerRolloutBuffer, RolloutBuffer  # noqa: E402
from upkie_b200.sim import UpkieSim  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, T = 4096, 8
    model = Model.standard_upkie()
    cfg = _abi.default_sim_config()
    sims = [UpkieSim(n, model=model, config=cfg, device=local) for _ in range(5)]
    for s in sims:
        s.set_autoreset(0, 0, rank * n)
        s.reset(seed=100 + rank, env_offset=rank * n)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5 + rank)
    tau = torch.tensor(model.tau_max, dtype=torch.float32, device=dev)
    local_buf = RolloutBuffer(T, n, 18, dev, compact=True)
    peer = PeerRolloutBuffer(T, n, 18, dev, compact=True)   # multicast stores
    peer2 = PeerRolloutBuffer(T, n, 18, dev, compact=True)  # peer stores
    peer3 = PeerRolloutBuffer(T, n, 18, dev, compact=True)  # deferred multicast push
    peer4 = PeerRolloutBuffer(T, n, 18, dev, compact=True)  # deferred peer stores
    pend3 = pend4 = None
    mc = peer.multicast_supported
    if not mc:
        print(f"rank {rank}: symmetric memory reports no multicast support on this box", flush=True)
    for t in range(T):
        a = torch.zeros((n, 6, 6), device=dev)
        a[:, :, 0] = float("nan")
        a[:, :, 5] = tau
        a[:, :, 2] = (torch.rand((n, 6), device=dev, generator=gen) * 2 - 1) * tau
        o, _, te, _ = local_buf.slot(t)
        sims[0].step_servos_compact(a, obs=o, terminated=te)
        if mc:
            sims[1].step_servos_multicast(a, *peer.multicast_slot(t))
        sims[2].step_servos_peers(a, *peer2.peer_slots(t))
        if mc:
            sims[3].step_servos_push(a, *peer3.local_slot(t), pend3)
            pend3 = peer3.push_descriptor(t, multicast=True)
        sims[4].step_servos_push(a, *peer4.local_slot(t), pend4)
        pend4 = peer4.push_descriptor(t, multicast=False)
    if mc:
        sims[3].push_rows(pend3)
        peer.publish()
        peer3.publish()
    sims[4].push_rows(pend4)
    peer2.publish()
    peer4.publish()
    torch.cuda.synchronize()
    # (1) transport: what every rank holds for rank r must be, byte for byte, what rank r holds for itself (its own
    # region, exchanged with a plain NCCL all-gather for the comparison). (2) physics: the TILE=2 instantiations are
    # compiled separately from the TILE=1 kernel that wrote `local_buf`; their rows agree to fp32 round-off, not
    # necessarily bit for bit (the compiler schedules / contracts each instantiation on its own).
    ref_rows = local_buf.obs.reshape(T, n, 18)
    ok = True
    for name, buf in (("multicast", peer if mc else None), ("peerstore", peer2),
                      ("deferred multicast", peer3 if mc else None), ("deferred peerstore", peer4)):
        if buf is None:
            continue
        own = buf.raw.clone()
        expect = torch.empty(world * buf.nbytes, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(expect, own)
        got = buf.gathered()
        good = torch.equal(got.reshape(-1), expect)
        rows = buf.obs.reshape(T, n, 18)
        close = bool(torch.allclose(rows, ref_rows, rtol=0, atol=2e-2)) and float((rows - ref_rows).abs().median()) < 1e-6
        flags = bool(torch.equal(buf.terminated, local_buf.terminated))
        ok = ok and good and close and flags
        print(f"rank {rank}: {name} rollout {'MATCHES' if good else 'DIFFERS FROM'} every rank's own copy "
              f"({world} ranks x {buf.nbytes} bytes); rows vs the TILE=1 kernel: max |d| "
              f"{float((rows - ref_rows).abs().max()):.2e}, terminated {'equal' if flags else 'DIFFER'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
# SPDX-License-Identifier: Apache-2.0
"""Benchmark of the vectorised Upkie env-step path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--workload servos|pendulum|mpc]
    python bench.py --impl reference ...      # CPU arm (oracle port on host cores)

A "step" is one pass of the hot path over one batch: one 5 ms control tick
(5 x 1 ms physics substeps) of every env of the batch = one kernel launch.
Prints ONE JSON line on rank 0.

Workloads (BASELINE.json configs):
  servos   (default) configs[2]/[4]: 65536 UpkieServos envs per GPU, pure torque
           actions ~ U(-tau_max, tau_max), floor friction ~ U(0.5, 1.2), initial
           pitch ~ U(-0.3, 0.3), link inertias x (1 + U(-0.2, 0.2)), fall/height
           termination with fused next-step auto-reset; env-index sharded across
           GPUs (weak scaling), NCCL all-gather of the [T=32] rollout buffer.
  pendulum configs[1]: 4096 ground-velocity envs, actions ~ U(-3, 3) m/s.
  mpc      configs[3]: 4096 robots x horizon-16 box-QP per tick.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per env-step (DESIGN.md "Roofline")
B_ALG = {"servos": 542 + 284 + 3 * 4 * 2, "pendulum": 346 + 3 * 4 * 2 + 16, "mpc": 157}
# compact rollout records: observation rows 72 B instead of 120 B, no reward (4 B) / truncated (1 B) stores
B_ALG_SERVOS_COMPACT = B_ALG["servos"] - 48 - 5
N_ACTION_BUFFERS = 16
ROLLOUT_T = 32


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = None

    def _run_nvml(self):
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {
            "hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else nv.nvmlClocksThrottleReasonHwSlowdown,
            "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
            "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
            "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap,
        }
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            try:
                power = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                power = 0.0
            self.samples.append(
                [str(sm), str(mx), str(power)] + ["Active" if (reasons & bits[k]) else "Not Active" for k in names]
            )
            self._stop.wait(0.01)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for s in self.samples for k in range(4) if len(s) > 3 + k and s[3 + k] == "Active"})
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons,
            "samples": len(self.samples),
        }


# ---- CPU arm ---------------------------------------------------------------------------------

def servos_config():
    from upkie_b200 import _abi

    cfg = _abi.default_sim_config()
    cfg.servos_fall_termination = 1
    cfg.min_base_height = 0.15
    cfg.rand_pitch = 0.3
    if os.environ.get("UPKIE_BENCH_PGS_TOL"):  # developer knob (profiles/r01_variants.md)
        cfg.pgs_tolerance = float(os.environ["UPKIE_BENCH_PGS_TOL"])
    return cfg


def cpu_servos(n_envs, steps, threads, seed=2025):
    """Times the oracle (CPU restatement) on a bounded sample of the servos workload."""
    from oracle import oracle
    from upkie_b200.model import Model

    oracle.build()
    model = Model.standard_upkie()
    cfg = servos_config()
    rng = np.random.default_rng(seed)
    sim = oracle.OracleSim(model, cfg, n_envs, threads=threads)
    sim.set_randomization(friction=rng.uniform(0.5, 1.2, n_envs), inertia_eps=rng.uniform(-0.2, 0.2, (n_envs, 6)))
    init = np.zeros((n_envs, 25))
    init[:, 2] = 0.6
    pitch = rng.uniform(-0.3, 0.3, n_envs)
    init[:, 3] = np.cos(pitch / 2)
    init[:, 5] = np.sin(pitch / 2)
    sim.reset(init)
    tau = np.asarray(model.tau_max)
    act = np.zeros((n_envs, 6, 6))
    act[:, :, 0] = np.nan
    act[:, :, 5] = tau
    act[:, :, 2] = rng.uniform(-1, 1, (n_envs, 6)) * tau
    sim.step_servos(act)  # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        act[:, :, 2] = rng.uniform(-1, 1, (n_envs, 6)) * tau
        _, _, term, _ = sim.step_servos(act)
        if term.any():
            sim.reset(init, mask=term)
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path. The
    reference itself (pybullet + gymnasium + upkie_description) cannot be
    installed here (DESIGN.md "Reference arm"), so this times the oracle port
    with all host threads."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample per step: calibrate so that the whole run ends within ~2 minutes
    rate_est, _ = cpu_servos(4096, 2, cores)
    budget_s = 100.0 / max(1, args.steps + args.warmup)
    n_sample = int(min(16384, max(256, rate_est * budget_s / 4)))
    n_sample = 1 << (n_sample.bit_length() - 1)
    rates = []
    for _ in range(args.warmup):
        cpu_servos(n_sample, 4, cores)
    t_total = 0.0
    for _ in range(args.steps):
        r, dt = cpu_servos(n_sample, 4, cores)
        rates.append(r)
        t_total += dt
    value = float(np.mean(rates))
    line = {
        "impl": "reference",
        "metric": "env-steps/sec",
        "value": value,
        "unit": "env-steps/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(1, args.steps),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": "UpkieServos 6-DoF torque actions, domain-randomized (BASELINE configs[2]/[4]); CPU port, "
            f"each step = {n_sample} envs x 4 ticks sample",
            "envs_per_step": n_sample,
        },
        "cpu_baseline": {
            "value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n_sample} envs x 4 ticks per step, {args.steps} steps, oracle fp64, {cores} threads",
        },
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm -------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="servos", choices=["servos", "pendulum", "mpc", "plumbing"])
    ap.add_argument("--envs-per-gpu", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from upkie_b200 import _abi
    from upkie_b200.envs import B200VectorEnv
    from upkie_b200.model import Model
    from upkie_b200.robot_state import RobotState, RobotStateRandomization

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # host buffers and the stepping thread on the GPU's NUMA node (the e2e path is PCIe-bound)
    from upkie_b200.numa import bind_to_gpu_node, gpu_numa_node

    previous_affinity = bind_to_gpu_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The rollout all-gather runs on NCCL's own stream while the next rollout simulates. The step kernel
        # occupies every SM (255 registers x 224 threads leave no room for a second block), so the collective's
        # CTAs only get SMs when a simulation block retires. Measured on 2 GPUs (tools/run_2gpu_variants.sh): a
        # high-priority NCCL stream or fewer CTAs (NCCL_MAX_CTAS) make it worse, the default is best.
        pg_options = None
        if os.environ.get("UPKIE_BENCH_NCCL_PRIORITY", "0") == "1":
            pg_options = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=pg_options)

    model = Model.standard_upkie()
    peaks, peaks_kind = read_peaks()
    W = max(3, args.warmup)
    K = args.steps

    if args.workload == "plumbing":
        print(json.dumps(bench_plumbing(torch, dev, model)), flush=True)
        return
    if args.workload == "mpc":
        result = bench_mpc(args, torch, dev, rank, world, K, W)
    else:
        result = bench_env(args, torch, dist, dev, rank, world, model, K, W)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    units, t_ms, kernel_ms, e2e, launches, clocks, config, n_per_gpu = result
    if previous_affinity is not None:
        os.sched_setaffinity(0, previous_affinity)  # the CPU baseline below uses every host core
    config["host_numa"] = (f"stepping thread and pinned buffers on NUMA node {gpu_numa_node(local_rank)} of the GPU"
                           if previous_affinity is not None else "no NUMA binding (single node or unknown topology)")
    total_units = units * world
    value = total_units / (t_ms * 1e-3)
    b_alg = B_ALG_SERVOS_COMPACT if config.get("rollout_record", "").startswith("compact") else B_ALG[args.workload]
    achieved = b_alg * n_per_gpu / (kernel_ms * 1e-3) / 1e9  # GB/s per GPU, dominant kernel
    line = {
        "metric": "env-steps/sec" if args.workload != "mpc" else "qp-solves/sec",
        "value": value,
        "unit": "env-steps/s" if args.workload != "mpc" else "qp-solves/s",
        "n_gpus": world,
        "steps": K,
        "warmup": W,
        "ms_per_step": t_ms / K,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": config,
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "roofline": {
            "bound": "hbm",
            "achieved": achieved,
            "peak": peaks["hbm_gbs"],
            "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"],
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel (ncu --set full,
            # profiles/r01_ncu_summary.md): below the algorithmic bytes because the 12 MB of robot state stay in the
            # 126 MB L2 between launches (reads = actions + state + randomisation, writes almost nil)
            "traffic": 22_959_360 if (args.workload == "servos" and n_per_gpu == 65536) else None,
            "peak_kind": f"{peaks_kind} (MEASURED_PEAKS.json hbm_gbs)" if peaks_kind == "measured" else "fallback 6650 GB/s",
            "algorithmic_bytes_per_unit": b_alg,
            "kernel_ms": kernel_ms,
            "note": "fp32 issue-bound path (DESIGN.md): HBM fraction is reported as asked, the binding bound is the "
                    "fp32 pipe; see fp32_issue",
        },
    }
    if args.workload != "mpc":
        # secondary roofline: non-tensor fp32 issue slots
        sm_mhz = clocks.get("sm_mhz") or 1700.0
        # warp instructions per env-step of the servos workload, ncu smsp__inst_executed.sum / warps
        # (profiles/r01_ncu_summary.md, paired f32x2 legs, end of round 1); 78.5 % of them FFMA(2)/FMUL(2)/FADD(2)
        instr_per_env_step = 13_921
        sched_cycles = 148 * 4 * sm_mhz * 1e6  # issue slots per second (one warp instruction each)
        ipc = instr_per_env_step * (n_per_gpu / 32.0) / (kernel_ms * 1e-3) / sched_cycles
        line["roofline"]["fp32_issue"] = {
            "ipc_per_scheduler": ipc,
            "peak_ipc": 1.0,
            "frac": ipc,
            # tools/micro/ffma2_bench.cu on this pool: three-register scalar FFMA saturates at 0.59 inst/cycle/scheduler
            "measured_scalar_ffma_ceiling_ipc": 0.59,
            "fp_instr_frac_of_ceiling": 0.785 * ipc / 0.59,
            "instr_per_env_step": instr_per_env_step,
            "exact_for": "servos workload (the pendulum front-end changes the count by < 2 %)",
        }
    if world == 1 and not args.no_cpu_baseline and args.workload != "mpc":
        cores = os.cpu_count() or 1
        n_cpu = 16384
        rate, dt = cpu_servos(n_cpu, 8, cores)
        line["cpu_baseline"] = {
            "value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n_cpu} envs x 8 ticks of the same workload, oracle fp64 ({dt:.1f} s wall, {cores} threads)",
        }
    elif world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_mpc_baseline()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_env(args, torch, dist, dev, rank, world, model, K, W):
    from upkie_b200 import _abi
    from upkie_b200.envs import B200VectorEnv
    from upkie_b200.robot_state import RobotState, RobotStateRandomization

    servos = args.workload == "servos"
    n = args.envs_per_gpu or (65536 if servos else 4096)
    gen = torch.Generator(device=dev)
    gen.manual_seed(2025 + rank)
    if servos:
        cfg = servos_config()
        env = B200VectorEnv(n, "servos", config=cfg, device=dev.index, autoreset_mode="next_step",
                            env_offset=rank * n, model=model)
        mu = torch.empty(n, device=dev).uniform_(0.5, 1.2, generator=gen)
        eps = torch.empty((n, 6), device=dev).uniform_(-0.2, 0.2, generator=gen)
        env.sim.set_randomization(friction=mu, inertia_eps=eps)
        tau = torch.tensor(model.tau_max, dtype=torch.float32, device=dev)
        acts = []
        for _ in range(N_ACTION_BUFFERS):
            a = torch.zeros((n, 6, 6), device=dev)
            a[:, :, 0] = float("nan")
            a[:, :, 5] = tau
            a[:, :, 2] = (torch.rand((n, 6), device=dev, generator=gen) * 2 - 1) * tau
            acts.append(a.contiguous())
        # rollout records: "compact" = position / velocity / torque rows + terminated (73 B/env/step); the constants
        # of the reference (temperature, voltage, reward, truncated) are not written nor gathered. "full" = 126 B.
        compact_rollout = os.environ.get("UPKIE_BENCH_ROLLOUT", "compact") == "compact"
        if compact_rollout:
            def step(a, obs=None, reward=None, terminated=None, truncated=None):
                return env.sim.step_servos_compact(a, obs=obs, terminated=terminated)
        else:
            step = env.sim.step_servos
        obs_bytes = (18 if compact_rollout else 30) * 4
        act_bytes = 36 * 4
    else:
        init = RobotState(randomization=RobotStateRandomization(pitch=0.1))
        env = B200VectorEnv(n, "pendulum", device=dev.index, autoreset_mode="next_step", env_offset=rank * n,
                            model=model, init_state=init)
        acts = [((torch.rand((n, 1), device=dev, generator=gen) * 2 - 1) * 3.0).contiguous()
                for _ in range(N_ACTION_BUFFERS)]
        step = env.sim.step_pendulum
        compact_rollout = False
        obs_bytes = 4 * 4
        act_bytes = 4
    env.sim.set_autoreset(1, 2025, rank * n)
    env.sim.reset(seed=2025, env_offset=rank * n)

    # rollout buffer gathered over NVLink once per T steps
    from upkie_b200.sharding import PeerRolloutBuffer, RolloutBuffer

    # two buffers: the gather of rollout r (NVLink) overlaps the simulation of r + 1. "peer": symmetric-memory
    # buffers, every rank pushes its slot to the peers with the copy engines (no SM); "nccl": all_gather_into_tensor
    # Measured (tools/run_2gpu_variants.sh, tools/run_8gpu.sh): 2 GPUs peer 96 % vs nccl 84 % weak-scaling efficiency;
    # 8 GPUs nccl 64 %, the first (unstaggered, one-stream) peer push collapsed there -> nccl stays the default
    # beyond 2 GPUs until the staggered push is validated at 8.
    gather_mode = os.environ.get("UPKIE_BENCH_GATHER", "peer" if world == 2 else "nccl") if world > 1 else "none"
    if gather_mode in ("peer", "multicast"):
        try:
            rollouts = [PeerRolloutBuffer(ROLLOUT_T, n, obs_bytes // 4, dev, compact=compact_rollout) for _ in range(2)]
            if gather_mode == "multicast" and not (servos and compact_rollout and rollouts[0].multicast_supported):
                raise RuntimeError("NVSwitch multicast needs the compact servos records and multicast-capable symmetric memory")
        except Exception as exc:  # symmetric memory unavailable on this box: fall back to NCCL's collective
            print(f"bench.py: symmetric-memory rollout buffer unavailable ({exc!r}); using NCCL all-gather", file=sys.stderr)
            gather_mode = "nccl"
    if gather_mode not in ("peer", "multicast"):
        rollouts = [RolloutBuffer(ROLLOUT_T, n, obs_bytes // 4, dev, compact=compact_rollout) for _ in range(2)]
    works = [None, None]

    for k in range(W):
        step(acts[k % N_ACTION_BUFFERS])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = env.sim.launches
    events = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    end = torch.cuda.Event(enable_timing=True)
    # ncu --profile-from-start off: "1" brackets the timed device loop, "e2e" the host-buffer loop
    profiling = os.environ.get("UPKIE_BENCH_CUDA_PROFILER", "") not in ("", "e2e")
    profiling_e2e = os.environ.get("UPKIE_BENCH_CUDA_PROFILER", "") == "e2e"
    with ClockSampler(dev.index) as clk:
        torch.cuda.synchronize()
        if profiling:
            torch.cuda.profiler.start()
        events[0].record()
        for k in range(K):
            # the kernel writes observation / reward / masks straight into the rollout slot of this step
            cur = (k // ROLLOUT_T) % 2
            if k % ROLLOUT_T == 0 and works[cur] is not None:
                # the gather that last read this buffer must be done before it is overwritten
                if gather_mode == "multicast":
                    pass  # stream order: publish() already ran on this stream
                elif gather_mode == "peer":
                    rollouts[cur].wait()
                else:
                    works[cur].wait()
                works[cur] = None
            if gather_mode == "multicast":
                # EXPERIMENTAL (UPKIE_BENCH_GATHER=multicast): the kernel's rows go to the NVSwitch multicast address
                # of this rank's slot and land in every GPU's buffer; a barrier per rollout replaces the gather
                env.sim.step_servos_multicast(acts[k % N_ACTION_BUFFERS], *rollouts[cur].multicast_slot(k))
            else:
                so, sr, ste, stru = rollouts[cur].slot(k)
                step(acts[k % N_ACTION_BUFFERS], obs=so, reward=sr, terminated=ste, truncated=stru)
            events[k + 1].record()
            if world > 1 and (k + 1) % ROLLOUT_T == 0:
                # one gather of the [T, n, 126 B] buffer per rollout, asynchronous
                if gather_mode == "multicast":
                    rollouts[cur].publish()
                elif gather_mode == "peer":
                    works[cur] = rollouts[cur].push()
                else:
                    _, works[cur] = rollouts[cur].gather_raw(async_op=True)
        for i_, w_ in enumerate(works):
            if w_ is not None:
                if gather_mode == "multicast":
                    pass
                elif gather_mode == "peer":
                    rollouts[i_].wait()
                else:
                    w_.wait()
        end.record()  # after the last step / all-gather queued on this stream
        torch.cuda.synchronize()
        if profiling:
            torch.cuda.profiler.stop()
        if world > 1:
            dist.barrier()
    total_ms = events[0].elapsed_time(end)
    per_step = np.array([events[k].elapsed_time(events[k + 1]) for k in range(K)])
    kernel_ms = float(np.median(per_step)) if world == 1 else float(np.min(per_step))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_ms = float(t.item()) / K
    launches = env.sim.launches - launches0

    # e2e through the public VectorEnv API with HOST buffers (H2D + kernel + D2H per step)
    # this step's inputs live in pinned host memory (4 rotating buffers), outputs land in pinned memory
    host_acts = [a.cpu().pin_memory().numpy() for a in acts[:4]]
    Ke = max(10, min(K, 400))
    for k in range(3):
        env.step(host_acts[k % 4])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if profiling_e2e:
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    for k in range(Ke):
        env.step(host_acts[k % 4])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if profiling_e2e:
        torch.cuda.profiler.stop()
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {
        "value": n * world * Ke / float(te.item()),
        "unit": "env-steps/s",
        "h2d_bytes_per_step": n * act_bytes,
        # servos: position/velocity/torque rows (72 B) + terminated; temperature, voltage, reward and truncated
        # are constants of the reference that the env fills once on the host (host path)
        "d2h_bytes_per_step": n * ((72 if servos else obs_bytes) + 1),
        "steps": Ke,
        "api": "B200VectorEnv.step(numpy action) -> numpy obs/reward/terminated/truncated",
    }
    config = {
        "workload": (
            "UpkieServos 6-DoF torque actions, 65536 envs/GPU, friction~U(0.5,1.2), init pitch~U(+-0.3), inertia "
            "eps~U(+-0.2), fall/height termination + fused next-step autoreset (BASELINE configs[2], per-GPU shard "
            "of configs[4])"
            if servos else "UpkiePendulum (UpkieGroundVelocity) 4096 envs, actions~U(-3,3) m/s (BASELINE configs[1])"
        ),
        "envs_per_gpu": n,
        "global_envs": n * world,
        "rollout_record": ("compact 73 B/env/step (position, velocity, torque rows + terminated; the reference's constants "
                           "temperature, voltage, reward, truncated are not stored)" if compact_rollout
                           else f"{obs_bytes + 6} B/env/step"),
        "substeps_per_step": 5,
        "joint_limit_rows": bool(getattr(env.config, "joint_limits", 0)),  # Bullet's hip / knee limit constraints: off
        # by default in round 1; on this torque workload they would be active on ~20 % of the robot-ticks
        "parallelism": f"env-index sharded x{world}" + (
            ("; rollout buffer [32] pushed to the peers' symmetric-memory buffers by the copy engines over NVLink"
             if gather_mode == "peer" else
             "; rollout rows stored to the NVSwitch multicast address of the symmetric buffer (experimental)"
             if gather_mode == "multicast" else "; NCCL all-gather of the [32] rollout buffer") if world > 1 else ""),
        "l2": f"{N_ACTION_BUFFERS} rotating action buffers ({N_ACTION_BUFFERS * n * act_bytes / 1e6:.0f} MB"
              " vs 126 MB L2); robot state stays resident by design",
    }
    return n * K, t_ms * K, kernel_ms, e2e, launches, clk.summary(), config, n


def bench_plumbing(torch, dev, model, steps=10_000):
    """BASELINE configs[0]: ONE Upkie-PyBullet-Pendulum-equivalent env at 200 Hz under the README PD policy
    (README.md:62-64), 10 k steps, reset on `terminated`, through the public env API with host arrays; the same
    loop on the CPU oracle beside it (single thread)."""
    from oracle import oracle
    from upkie_b200 import _abi
    from upkie_b200.envs import B200VectorEnv

    gains = np.array([10.0, 1.0, 0.0, 0.1], dtype=np.float32)
    env = B200VectorEnv(1, "pendulum", model=model, device=dev.index)
    obs, _ = env.reset(seed=0)
    for _ in range(50):
        obs, _, term, _, _ = env.step((gains @ obs[0]).reshape(1, 1))
    t0 = time.perf_counter()
    resets = 0
    for _ in range(steps):
        obs, _, term, _, _ = env.step((gains @ obs[0]).reshape(1, 1))
        if term[0]:
            obs, _ = env.reset()
            resets += 1
    gpu_rate = steps / (time.perf_counter() - t0)
    pitch_final = float(obs[0, 0])
    cfg = _abi.default_sim_config()
    osim = oracle.OracleSim(model, cfg, 1)
    init = np.zeros((1, 25))
    init[0, 2], init[0, 3] = 0.6, 1.0
    osim.reset(init)
    o = osim.reset_obs(4)
    t0 = time.perf_counter()
    for _ in range(steps):
        o, _, oterm, _ = osim.step_gyropod((gains.astype(np.float64) @ o[0]).reshape(1, 1), 1)
        if oterm[0]:
            osim.reset(init)
            o = osim.reset_obs(4)
    cpu_rate = steps / (time.perf_counter() - t0)
    return {
        "metric": "env-steps/sec", "value": gpu_rate, "unit": "env-steps/s", "n_gpus": 1, "steps": steps, "warmup": 50,
        "ms_per_step": 1e3 / gpu_rate, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "single UpkiePendulum env, 200 Hz, README PD policy, 10k steps (BASELINE configs[0]); "
                               "latency-bound plumbing case, not the throughput configuration",
                   "resets": resets, "final_pitch": pitch_final},
        "e2e": {"value": gpu_rate, "unit": "env-steps/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 22,
                "api": "B200VectorEnv(1, 'pendulum').step(numpy)"},
        "gpu_launches": steps,
        "cpu_baseline": {"value": cpu_rate, "unit": "env-steps/s", "cores": 1, "kind": "port",
                         "sample": f"{steps} steps of the same closed loop on the fp64 oracle, 1 thread"},
    }


def bench_mpc(args, torch, dev, rank, world, K, W):
    from upkie_b200 import _abi
    from upkie_b200.mpc import BatchedMPCBalancer

    n = args.envs_per_gpu or 4096
    cfg = _abi.default_mpc_config()
    cfg.nb_timesteps = 16
    mpc = BatchedMPCBalancer(n, config=cfg, device=dev.index)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4 + rank)

    def U(lo, hi, *shape):
        return torch.rand(shape, device=dev, generator=gen) * (hi - lo) + lo

    xs = [torch.stack([U(-0.5, 0.5, n), U(-0.2, 0.2, n), U(-0.5, 0.5, n), U(-1, 1, n)], dim=1).contiguous()
          for _ in range(N_ACTION_BUFFERS)]
    vt = U(-1, 1, n)
    contact = torch.ones(n, dtype=torch.uint8, device=dev)
    for k in range(W):
        mpc.step_tensors(xs[k % N_ACTION_BUFFERS], vt, contact, 0.005)
    torch.cuda.synchronize()
    events = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    with ClockSampler(dev.index) as clk:
        events[0].record()
        for k in range(K):
            mpc.step_tensors(xs[k % N_ACTION_BUFFERS], vt, contact, 0.005)
            events[k + 1].record()
        torch.cuda.synchronize()
    total_ms = events[0].elapsed_time(events[K])
    per_step = np.array([events[k].elapsed_time(events[k + 1]) for k in range(K)])
    xh = [x.cpu().numpy() for x in xs[:4]]
    vth, ch = vt.cpu().numpy(), contact.cpu().numpy()
    Ke = max(10, min(K, 400))
    t0 = time.perf_counter()
    for k in range(Ke):
        mpc.step(xh[k % 4], vth, ch, 0.005)
    e2e_s = time.perf_counter() - t0
    e2e = {"value": n * Ke / e2e_s, "unit": "qp-solves/s", "h2d_bytes_per_step": n * (16 + 4 + 1),
           "d2h_bytes_per_step": n * 4, "steps": Ke, "api": "BatchedMPCBalancer.step(numpy) -> numpy"}
    config = {"workload": "MPC balancer 4096 robots x horizon-16 box-QP per 5 ms tick (BASELINE configs[3])",
              "robots": n, "horizon": 16, "l2": "working set < L2 by nature (4096 x 157 B)"}
    return n * K, total_ms, float(np.median(per_step)), e2e, K, clk.summary(), config, n


def cpu_mpc_baseline():
    from oracle import oracle
    from upkie_b200 import _abi

    cfg = _abi.default_mpc_config()
    cfg.nb_timesteps = 16
    m = oracle.OracleMpc(cfg)
    rng = np.random.default_rng(0)
    n = 4096
    x0 = np.stack([rng.uniform(-0.5, 0.5, n), rng.uniform(-0.2, 0.2, n), rng.uniform(-0.5, 0.5, n), rng.uniform(-1, 1, n)], 1)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps):
        m.step(x0, rng.uniform(-1, 1, n), np.ones(n, np.uint8), 0.005, np.zeros(n), threads=cores)
    dt = time.perf_counter() - t0
    return {"value": n * reps / dt, "unit": "qp-solves/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x 4096 solves, fp64 dense active-set oracle, {cores} threads"}


if __name__ == "__main__":
    main()

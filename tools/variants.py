#!/usr/bin/env python
"""Build-time variant explorer for the step kernel (developer tool).

    python tools/variants.py build      # here (no GPU): compiles variants/*.so
    python tools/variants.py run        # on the GPU box: times each variant
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "upkie_b200", "csrc")
OUT = os.path.join(ROOT, "variants")
BASE = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
        "-Xcompiler", "-fPIC", "--use_fast_math"]
VARIANTS = {
    "paired_s1": [],
    "paired_s0": ["-DUPKIE_PHASE_SYNC_LEVEL=0"],
    "scalar_s1": ["-DUPKIE_PAIRED_LEGS=0"],
    "paired_s1_r128": ["-DUPKIE_MAX_THREADS=512", "-DUPKIE_MIN_BLOCKS=1"],  # run with UPKIE_B200_BLOCK=448
}
SOURCES = ["upkie_b200.cu", "step_device.cu", "step_host.cu"]


def build():
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        cmd = BASE + flags + ["-Xptxas", "-v", "-o", os.path.join(OUT, f"lib_{name}.so")] + [os.path.join(CSRC, s) for s in SOURCES]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        lines = out.splitlines()
        info = ""
        for i, l in enumerate(lines):
            if "k_stepILi0ELi1ELi0ELi0E" in l:
                info = " | ".join(x.strip() for x in lines[i + 1:i + 3])
        print(f"{name:16s} rc={p.returncode} {info}")


def run():
    names = sys.argv[2:] or list(VARIANTS)
    for name in names:
      try:
        env = dict(os.environ, UPKIE_B200_LIB=os.path.join(OUT, f"lib_{name}.so"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "300", "--warmup", "20",
                            "--no-cpu-baseline"], env=env, capture_output=True, text=True, timeout=150)
        import json
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"{name:16s} value={j['value']:.4e} kernel_ms={j['roofline']['kernel_ms']:.4f} e2e={j['e2e']['value']:.3e}", flush=True)
        except Exception:
            print(name, "FAILED", r.stdout[-300:], r.stderr[-600:], flush=True)
      except subprocess.TimeoutExpired:
        print(name, "TIMEOUT", flush=True)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()

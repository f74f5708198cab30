#!/usr/bin/env python3
# SPDX-License-Identifier: Apache-2.0
"""Static SASS instruction counts of one step kernel, bucketed by the source function each instruction's line
belongs to (needs the library built with -lineinfo, which build.py does). Developer tool, runs without a GPU:

    python tools/static_breakdown.py [kernel-substring]      # default: k_stepILi0ELi1ELi0ELi1E (bench kernel)

The "substep:" buckets (and `servo_substep`) are inside the `nb_substeps` loop; the PGS bucket holds two inlined
copies of the six-row sweep, each run once per pair of sweeps.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "upkie_b200", "libupkie_b200.so")
CSRC = os.path.join(ROOT, "upkie_b200", "csrc")


def function_starts(path):
    """[(line, name)] of function definitions / named lambdas in a source file (good enough for bucketing)."""
    out = []
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:UPKIE_HD|__global__|static|inline|__device__)[^;=]*?\b(\w+)\s*\(")
    lam = re.compile(r"^\s*auto\s+(\w+)\s*=\s*\[")
    for i, line in enumerate(open(path), 1):
        m = pat.match(line) or lam.match(line)
        if m and m.group(1) not in ("if", "for", "while", "return"):
            out.append((i, m.group(1)))
    return out


PHASE_MARKERS = [  # (regex on a source line of physics_substep_paired, phase that starts there)
    (r"float R\[9\];", "rotation, base inertia"),
    (r"legs_pass12\(P, S\.q", "ABA passes 1-2, both legs"),
    (r"ldl6\(IA0\);", "base LDL^T + solve"),
    (r"legs_pass3\(P, lc", "ABA pass 3"),
    (r"// gravity as a uniform frame acceleration", "velocity update"),
    (r"// -- collision detection", "collision detection"),
    (r"// contact directions in base coordinates", "contact Jacobians, wheel velocities"),
    (r"// Delassus matrix W = J M\^-1 J\^T without", "Delassus matrix"),
    (r"// right-hand sides \(btMultiBodyConstraintSolver", "row setup"),
    (r"auto sweep = \[&\]", "PGS sweeps"),
    (r"for \(int it = 0; it < P\.pgs_iterations; it \+= 2\)", "PGS sweeps (two inlined copies of the six-row sweep, plain and with exit test) + loop control"),
    (r"// apply the total wheel impulses", "apply impulses"),
    (r"if \(slow\) limit_contact_solve", "joint-limit slow path"),
    (r"// -- position integration with the new velocities", "position integration"),
]


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else "k_stepILi0ELi1ELi0ELi1E"
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        text = None
        for cubin in sorted(os.listdir(tmp)):
            dis = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)],
                                 capture_output=True, text=True).stdout
            m = re.search(r"^\.text\.(\S*" + re.escape(want) + r"\S*):", dis, re.M)
            if m:
                start = m.end()
                nxt = re.search(r"^\.text\.\S+:", dis[start:], re.M)
                text = dis[start:start + nxt.start()] if nxt else dis[start:]
                print(f"kernel {m.group(1)[:90]}... in {cubin}")
                break
        if text is None:
            raise SystemExit(f"no kernel matching {want!r} in {LIB}")
    starts = {}

    def func_of(fn, ln):
        if fn not in starts:
            p = os.path.join(CSRC, fn)
            starts[fn] = function_starts(p) if os.path.exists(p) else []
        name = "?"
        for s_line, s_name in starts[fn]:
            if s_line <= ln:
                name = s_name
            else:
                break
        return name

    # phases of physics_substep_paired by line
    pair_src = open(os.path.join(CSRC, "sim_pair.cuh")).read().splitlines()
    fn_start = next(i for i, l in enumerate(pair_src, 1) if "void physics_substep_paired(" in l)
    phase_lines = []
    for rx, name in PHASE_MARKERS:
        for i in range(fn_start, len(pair_src) + 1):
            if re.search(rx, pair_src[i - 1]):
                phase_lines.append((i, name))
                break
    phase_lines.sort()

    def phase_of(ln):
        name = "substep prologue"
        for s_line, s_name in phase_lines:
            if s_line <= ln:
                name = s_name
        return name

    counts, packed = collections.Counter(), collections.Counter()
    chain, total = [], 0
    pending = []
    for line in text.splitlines():
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            if m.group(3):
                pending.append((os.path.basename(m.group(3)), int(m.group(4))))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", line):
            if pending:
                chain, pending = pending, []
            frames = [(fn, ln, func_of(fn, ln)) for fn, ln in chain]  # innermost first
            bucket = None
            for fn, ln, name in reversed(frames):  # outermost first
                if fn == "sim_pair.cuh" and ln >= fn_start:  # physics_substep_paired is the last function of the file
                    bucket = "substep: " + phase_of(ln)
                    break
            if bucket is None:
                named = [name for fn, ln, name in reversed(frames) if fn.endswith((".cuh", ".cu", ".h"))
                         and name not in ("step_env", "k_step", "__launch_bounds__", "prefetch", "?")]
                bucket = "tick: " + (named[0] if named else "step_env (front-end, observation, stores)")
            counts[bucket] += 1
            if re.search(r"\b(FFMA2|FMUL2|FADD2)\b", line):
                packed[bucket] += 1
            total += 1
    print(f"{total} static instructions ({sum(packed.values())} packed f32x2)")
    print(f"{'instr':>6s} {'share':>6s} {'f32x2':>6s}  bucket")
    for bucket, c in sorted(counts.items(), key=lambda kv: (not kv[0].startswith("substep"), -kv[1])):
        print(f"{c:6d} {100.0 * c / total:5.1f}% {packed[bucket]:6d}  {bucket}")


if __name__ == "__main__":
    main()

#!/bin/bash
# final build: the configs[1] / configs[3] lines of bench.py for the record
mkdir -p gpurun_out/r02l
timeout 200 python bench.py --workload pendulum --steps 1000 --warmup 100 > gpurun_out/r02l/bench_pendulum.json 2> gpurun_out/r02l/bench_pendulum.err
timeout 200 python bench.py --workload mpc --steps 300 --warmup 20 > gpurun_out/r02l/bench_mpc.json 2> gpurun_out/r02l/bench_mpc.err
python - <<'P'
import json
for f in ("bench_pendulum","bench_mpc"):
    try:
        j=json.loads(open(f"gpurun_out/r02l/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4f e2e %.4g kernel_ms %.4f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["kernel_ms"]), "cpu", (j.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
P

#!/bin/bash
mkdir -p gpurun_out/r02f
timeout 600 python -m pytest tests -m gpu -q -k "mpc or base_velocity" 2>&1 | tail -3
for H in 16 50; do
timeout 300 python bench.py --workload mpc --steps 300 --warmup 30 > gpurun_out/r02f/mpc.json 2> gpurun_out/r02f/mpc.err
done
tail -c 1500 gpurun_out/r02f/mpc.json
timeout 300 python bench.py --steps 30 --warmup 10 --no-cpu-baseline > gpurun_out/r02f/line.json 2> gpurun_out/r02f/line.err
python - <<'P'
import json
j=json.loads(open("gpurun_out/r02f/line.json").read().strip().splitlines()[-1])
for k,v in j.get("other_workloads",{}).items(): print(k, {kk:vv for kk,vv in v.items() if kk!="workload"} if isinstance(v,dict) else v)
P

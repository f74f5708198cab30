#!/bin/bash
# Round 2, session 2, GPU call 2: the whole GPU test-suite (body-contact kernels in their own instantiations), smoke,
# bench at steady state and at the driver's settings.
mkdir -p gpurun_out/r02i
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02i/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i/pytest_gpu.log
tail -8 gpurun_out/r02i/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02i/smoke.log 2>&1; tail -2 gpurun_out/r02i/smoke.log
timeout 600 python bench.py --steps 300 --warmup 20 > gpurun_out/r02i/bench_steady.json 2> gpurun_out/r02i/bench_steady.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-other-workloads > gpurun_out/r02i/bench_driver.json 2> gpurun_out/r02i/bench_driver.err
python - <<'P'
import json
for f in ("bench_steady","bench_driver"):
    try:
        j=json.loads(open(f"gpurun_out/r02i/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4f e2e %.4g kernel_ms %.4f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["kernel_ms"]))
        ow=j.get("other_workloads",{})
        for k,v in ow.items(): print("   ",k, {kk:vv for kk,vv in v.items() if kk!="workload"} if isinstance(v,dict) else v)
    except Exception as e: print(f, "ERR", e)
P

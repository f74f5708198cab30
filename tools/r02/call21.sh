#!/bin/bash
# Round 2, session 2: final profile pass of the benchmarked build (1 GPU): bench lines, steady / early `ncu --set full`
# captures of the step kernel, launch list, MPC capture, GPU test-suite, smoke.
mkdir -p gpurun_out/r02g
O=gpurun_out/r02g
timeout 600 python bench.py --steps 300 --warmup 20 > $O/bench_steady.json 2> $O/bench_steady.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err
UPKIE_BENCH_CUDA_PROFILER=1 timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --launch-skip 200 -c 1 -f -o $O/prof_steady python bench.py --steps 210 --warmup 20 --no-cpu-baseline --no-other-workloads > $O/ncu_steady.log 2>&1
UPKIE_BENCH_CUDA_PROFILER=1 timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --launch-skip 10 -c 1 -f -o $O/prof_early python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/ncu_early.log 2>&1
UPKIE_BENCH_CUDA_PROFILER=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mpc_step --launch-skip 60 -c 1 -f -o $O/prof_mpc python bench.py --workload mpc --steps 100 --warmup 10 --no-cpu-baseline > $O/ncu_mpc.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python - <<'P'
import json
for f in ("bench_steady","bench_driver","bench_reference"):
    try:
        j=json.loads(open(f"gpurun_out/r02g/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.4f e2e %.4g" % (j["value"], j.get("ms_per_step",0), j["e2e"]["value"]), "kernel_ms", (j.get("roofline") or {}).get("kernel_ms"), "cpu", (j.get("cpu_baseline") or {}).get("value"))
        for k,v in j.get("other_workloads",{}).items(): print("   ",k, {kk:vv for kk,vv in v.items() if kk!="workload"} if isinstance(v,dict) else v)
    except Exception as e: print(f, "ERR", e)
P
ls -la $O | tail -20

#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu9.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu9.log
tail -4 $O/pytest_gpu9.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench9_driver.json 2> $O/bench9_driver.err ) 2> $O/bench9_time.txt
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench9_reference.json 2> $O/bench9_reference.err ) 2>> $O/bench9_time.txt
timeout 300 python bench.py --workload pendulum --steps 1000 --warmup 100 --no-cpu-baseline > $O/bench9_pendulum.json 2> $O/bench9_pendulum.err
timeout 300 python bench.py --workload plumbing > $O/bench9_plumbing.json 2> $O/bench9_plumbing.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke9.txt 2>&1
cat $O/bench9_time.txt | grep real; cat $O/smoke9.txt | tail -2
python - <<PY
import json
d=json.loads(open("$O/bench9_driver.json").read().strip().splitlines()[-1])
print("driver", "%.4g"%d["value"], d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d["roofline"]["traffic"], d["roofline"].get("fp32_issue"))
print(d["other_workloads"])
print(d["cpu_baseline"])
d=json.loads(open("$O/bench9_reference.json").read().strip().splitlines()[-1]); print("ref", d["value"], d["ms_per_step"])
d=json.loads(open("$O/bench9_plumbing.json").read().strip().splitlines()[-1]); print("plumbing", d["value"], d["config"], d["cpu_baseline"]["value"])
PY
cat gpurun_out/parity_report.json | tail -16

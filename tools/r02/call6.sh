#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
python tools/variants.py run > $O/variants6.log 2>&1
cat $O/variants6.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu6.log
tail -4 $O/pytest_gpu6.log

#!/bin/bash
# lane = robot (packed f32x2 legs) against lane = leg (two lanes per robot, scalar legs + shuffles) on the ABA passes 1-2
mkdir -p gpurun_out/r02k
python -c "
import sys; sys.path.insert(0,'.')
from upkie_b200.model import Model; import bench
open('/tmp/pgs_model.bin','wb').write(bytes(Model.standard_upkie().to_struct()) + bytes(bench.servos_config()))"
timeout 120 tools/micro/lane_per_leg_bench | tee gpurun_out/r02k/lane_per_leg.txt

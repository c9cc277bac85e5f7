#!/bin/bash
# ncu capture of every level-0/1 launch class of one training step (tools/profile_blocks.py), sections without the
# heavy per-source counters so the report stays small; raw CSV exported on the box.
TAG=${1:-ncu}
O=gpurun_out/$TAG
mkdir -p $O
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --section ComputeWorkloadAnalysis --section InstructionStats"
timeout 400 ncu $SEC --clock-control none -c 40 -k regex:'dwconv|dw_wgrad|mlp_bwd_fused|mlp_fused|gn_dy|head_bwd' \
  -o /tmp/blocks python tools/profile_blocks.py > $O/ncu_blocks.log 2>&1
ncu -i /tmp/blocks.ncu-rep --page raw --csv > $O/blocks_raw.csv 2>/dev/null
ls -la /tmp/blocks.ncu-rep
SZ=$(stat -c %s /tmp/blocks.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -lt 30000000 ]; then cp /tmp/blocks.ncu-rep $O/; fi
(timeout 100 python tools/profile_blocks.py --time) > $O/blocks_time.log 2>&1
(timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-ops) > $O/bench_ops.json 2> $O/bench_ops.err
ls -la $O; du -sh gpurun_out

#!/bin/bash
# One GPU-box visit: parity tests, A/B bench variants (env toggles), per-block timing and an ncu capture.
# Everything lands in gpurun_out/.  Usage: bash tools/gpu_call.sh [tag]
TAG=${1:-call}
O=gpurun_out/$TAG
mkdir -p $O
(timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
(timeout 150 $B) > $O/bench_default.json 2> $O/bench_default.err
for v in "PCB_NO_DEEP_BWD=1" "PCB_NO_UP3=1" "PCB_UP3_XB=2" "PCB_BWD_OVERLAP=0"; do
  (timeout 150 env $v $B) > $O/bench_$v.json 2> $O/bench_$v.err
done
for f in $O/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['execution'])" 2>&1 | tail -1)"; done
(timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-ops) > $O/bench_ops.json 2> $O/bench_ops.err
(timeout 100 python tools/profile_blocks.py --time) > $O/blocks_time.log 2>&1
(timeout 100 python tools/time_fwd.py --split) > $O/time_fwd.log 2>&1
timeout 400 ncu --set full --clock-control none -c 80 -k regex:'dwconv|dw_wgrad|mlp_bwd_fused|mlp_fused|gn_dy|head_bwd' \
  -o $O/blocks python tools/profile_blocks.py > $O/ncu_blocks.log 2>&1
ncu -i $O/blocks.ncu-rep --page raw --csv > $O/blocks_raw.csv 2>/dev/null
ls -la $O

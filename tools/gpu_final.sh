#!/bin/bash
# Final visit of a round: parity tests, the bench lines, the launch list of one step and ncu captures of the top kernels.
TAG=${1:-final}
O=gpurun_out/$TAG
mkdir -p $O
(timeout 400 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -40) > $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
(timeout 200 python bench.py) > $O/bench_n1.json 2> $O/bench_n1.err
(timeout 150 python bench.py --mode infer --steps 3 --no-cpu-baseline) > $O/bench_infer_n1.json 2> $O/bench_infer_n1.err
for f in $O/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'][:40], d['roofline']['frac'])" 2>&1 | tail -1)"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file $O/launches_train_b4.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_launches.log 2>&1
bash tools/gpu_ncu_blocks.sh $TAG
FULL="ncu --set full --clock-control none --import-source on"
timeout 120 $FULL -k regex:mlp_bwd_fused --launch-skip 3 -c 1 -o $O/mlp_bwd_fused_l0 python tools/profile_blocks.py > $O/ncu_full1.log 2>&1
timeout 120 $FULL -k regex:dwconv_same_tiled -c 1 -o $O/dwconv_same_tiled_l0 python tools/profile_blocks.py > $O/ncu_full2.log 2>&1
timeout 120 $FULL -k regex:mlp_fused_kernel -c 1 -o $O/mlp_fused_l0 python tools/profile_blocks.py > $O/ncu_full3.log 2>&1
ls -la $O; du -sh gpurun_out

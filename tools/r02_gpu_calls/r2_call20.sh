#!/bin/bash
# round 2, call 20: the profiles the bench numbers are explained by — launch lists (train step, inference) and full ncu captures of the
# dominant kernels (source-level), final kernels
O=gpurun_out/r2c20
mkdir -p $O
(PCB_BWD_OVERLAP=0 timeout 600 python bench.py --config c2 --profile-ops --no-graph --steps 3 --warmup 3 --no-cpu-baseline --no-e2e) > $O/bench_ops.json 2> $O/bench_ops.err
grep -E "ms/step" $O/bench_ops.err > $O/train_ops.txt; head -12 $O/train_ops.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_train_b4.csv python bench.py --config c2 --steps 1 --warmup 0 --no-graph --no-cpu-baseline --no-e2e > $O/ncu_train.log 2>&1
tail -2 $O/ncu_train.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_infer_sb2.csv python bench.py --mode infer --volume 320 --sw-batch 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_infer.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'mlp_fused_kernel|mlp_bwd_ws_kernel|dwconv_same_tiled' -c 3 -o $O/l0_fwd python tools/profile_blocks.py --batch 2 > $O/ncu_l0a.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'mlp_bwd_ws_kernel|mlp_bwd_ws2_kernel' -c 3 -o $O/l0_bwd python tools/profile_blocks.py --batch 2 > $O/ncu_l0b.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'tn_gemm_ws_kernel|gemm_ws_kernel' -c 6 -o $O/deep python tools/profile_deep.py > $O/ncu_deep.log 2>&1
ls -la $O

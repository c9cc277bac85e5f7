#!/bin/bash
# round 2, call 5: source-level ncu capture of the level-0 kernels + inference sw_batch / path sweep
O=gpurun_out/r2c5
mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'mlp_fused_kernel|dwconv_same_tiled|mlp_bwd_ws_kernel|gemm_ws_kernel' -c 14 \
  -o $O/l0_kernels python tools/profile_blocks.py --batch 2 > $O/ncu.log 2>&1
ls -la $O
for sb in 2 4 8; do
  (timeout 300 python bench.py --mode infer --volume 480 --sw-batch $sb --steps 2 --no-cpu-baseline --no-e2e) > $O/infer_sb$sb.json 2> $O/infer_sb$sb.err
  python -c "
import json; d=json.load(open('$O/infer_sb$sb.json')); print('sw_batch $sb', round(d['value'],1), 'Mvox/s', round(d['ms_per_step'],1), 'ms; module path', round(d['execution']['module_path_ms_per_step'],1))" 2>&1 | tail -1
done

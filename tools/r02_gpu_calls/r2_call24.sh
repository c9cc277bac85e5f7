#!/bin/bash
# round 2, call 24: why did the resident 480^3 inference pass slow down?  A/B the tiled transposed stencil, repeat runs; TTA re-test
O=gpurun_out/r2c24
mkdir -p $O
(timeout 300 python -m pytest tests/test_tta_affinity.py tests/test_tta_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2)
for v in "X=1" "PCB_NO_UP3_TILED=1" "X=2"; do
  (env $v timeout 300 python bench.py --mode infer --volume 480 --sw-batch 2 --steps 3 --no-cpu-baseline) > $O/infer_$v.json 2> $O/infer_$v.err
  python -c "
import json; d=json.load(open('$O/infer_$v.json')); print('$v', round(d['value'],1), 'Mvox/s', round(d['ms_per_step'],1), 'ms; module', round(d['execution']['module_path_ms_per_step'],1), 'e2e', round(d['e2e']['value'],1))"
done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,temperature.gpu,power.draw --format=csv

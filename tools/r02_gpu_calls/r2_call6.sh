#!/bin/bash
# round 2, call 6: the driver's own command (pytest -m gpu over the whole tree), then inference sw_batch sweep with a launch list
O=gpurun_out/r2c6
mkdir -p $O
(timeout 1200 python -X faulthandler -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
for sb in 1 2 4; do
  (timeout 300 python bench.py --mode infer --volume 480 --sw-batch $sb --steps 2 --no-cpu-baseline --no-e2e) > $O/infer_sb$sb.json 2> $O/infer_sb$sb.err
  python -c "
import json; d=json.load(open('$O/infer_sb$sb.json')); print('sw_batch $sb', round(d['value'],1), 'Mvox/s', round(d['ms_per_step'],1), 'ms; module path', round(d['execution']['module_path_ms_per_step'],1))" 2>&1 | tail -1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_infer_sb2.csv python bench.py --mode infer --volume 320 --sw-batch 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_infer.log 2>&1
tail -3 $O/ncu_infer.log

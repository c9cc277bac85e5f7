#!/bin/bash
# round 2, call 13: column-split ws2 backward (NB = 1 shapes) — parity, A/B, c2 bench; MedNeXt-L config c4 on the GPU
O=gpurun_out/r2c13
mkdir -p $O
(timeout 600 python -X faulthandler -m pytest tests/test_mednext_bwd_gpu.py -m gpu -q -x --durations=3 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
for v in 0 1; do
  (PCB_BWD_SPLIT=$v timeout 300 python tools/time_bwd_ws.py --batch 2 --modes 2 2>&1 | tail -8 | sed "s/^/SPLIT=$v /") | tee $O/time_bwd_split$v.log
done
(timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline) > $O/bench_c2.json 2> $O/bench_c2.err
python -c "
import json; d=json.load(open('$O/bench_c2.json')); print('c2', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'roof', d['roofline']['kernel'][:40], d['roofline']['frac'])" 2>&1 | tail -1
(timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline) > $O/bench_c4.json 2> $O/bench_c4.err
tail -c 600 $O/bench_c4.err
python -c "
import json; d=json.load(open('$O/bench_c4.json')); print('c4', d['config']['workload'], round(d['value'],2), d['unit'], round(d['ms_per_step'],2), 'ms; e2e', d['e2e']['value'], 'roof', d['roofline']['kernel'][:40], d['roofline']['frac'], 'step', d['step_roofline'].get('frac'))" 2>&1 | tail -1

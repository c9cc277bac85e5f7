#!/bin/bash
# round 2, call 17: MedNeXt-L after the stage-ownership fix (sizing + BASELINE config c4), whole GPU suite
O=gpurun_out/r2c17
mkdir -p $O
for side in 96 224; do
  (PCB_DEBUG_HANG=150 timeout 200 python tools/time_train_step.py --size L --side $side --top 14 2>&1 | grep -v "^  File" | tail -18) | tee $O/time_L_$side.log
done
(timeout 900 python bench.py --config c4 --steps 3 --warmup 3) > $O/bench_c4.json 2> $O/bench_c4.err
tail -c 300 $O/bench_c4.err
python -c "
import json; d=json.load(open('$O/bench_c4.json')); print('c4', d['config']['workload'], round(d['value'],3), d['unit'], round(d['ms_per_step'],2), 'ms; e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['sample'][:120], 'roof', d['roofline']['kernel'][:40], d['roofline']['frac'], 'step', d['step_roofline'].get('frac'))" 2>&1 | tail -1
(timeout 1200 python -X faulthandler -m pytest tests -m gpu -q -x --durations=5 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log

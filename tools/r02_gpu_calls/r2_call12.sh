#!/bin/bash
# round 2, call 12: clean per-op table of the train step (weight gradients on the main stream), source-level ncu of tn_gemm_ws,
# and the BASELINE configs c3 / c4 on one GPU
O=gpurun_out/r2c12
mkdir -p $O
(PCB_BWD_OVERLAP=0 timeout 600 python bench.py --config c2 --profile-ops --no-graph --steps 3 --warmup 3 --no-cpu-baseline --no-e2e) > $O/bench_ops.json 2> $O/bench_ops.err
grep -E "ms/step" $O/bench_ops.err | head -60
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'tn_gemm_ws_kernel' -c 3 -o $O/tn_ws python tools/profile_deep.py > $O/ncu.log 2>&1
tail -2 $O/ncu.log
for c in c3 c4; do
  (timeout 900 python bench.py --config $c --steps 5 --warmup 3) > $O/bench_$c.json 2> $O/bench_$c.err
  python -c "
import json; d=json.load(open('$O/bench_$c.json')); print('$c', d['config']['workload'], round(d['value'],2), d['unit'], round(d['ms_per_step'],2), 'ms; e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'roof', d['roofline']['kernel'][:40], d['roofline']['frac'], 'step', d['step_roofline'].get('frac'))" 2>&1 | tail -1
done

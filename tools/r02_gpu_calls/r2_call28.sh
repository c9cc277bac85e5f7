#!/bin/bash
# round 2, call 28: exercise the inference re-measure branch of bench.py once (forced), 320^3 volume
O=gpurun_out/r2c28
mkdir -p $O
(PCB_BENCH_REMEASURE=1 timeout 120 python bench.py --mode infer --volume 320 --sw-batch 2 --steps 2 --no-cpu-baseline --no-e2e) > $O/infer.json 2> $O/infer.err
tail -c 300 $O/infer.err
python -c "
import json; d=json.load(open('$O/infer.json')); print(round(d['value'],1), round(d['ms_per_step'],1), d['execution']['remeasured'], d['execution']['module_path_ms_per_step'])"

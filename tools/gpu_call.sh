#!/bin/bash
# One GPU-box visit: parity tests, A/B bench variants (env toggles), infer bench, block-level ncu.  Output: gpurun_out/<tag>/
TAG=${1:-call}
O=gpurun_out/$TAG
mkdir -p $O
(timeout 400 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -60) > $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
(timeout 150 $B) > $O/bench_default.json 2> $O/bench_default.err
for v in $VARIANTS; do
  (timeout 150 env $v $B) > $O/bench_$v.json 2> $O/bench_$v.err
done
for f in $O/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['execution']['timed_region'])" 2>&1 | tail -1)"; done
(timeout 150 python bench.py --mode infer --steps 3 --no-cpu-baseline) > $O/infer_b2.json 2> $O/infer_b2.err
(timeout 150 python bench.py --mode infer --steps 3 --no-cpu-baseline --sw-batch 4) > $O/infer_b4.json 2> $O/infer_b4.err
for f in $O/infer_*.json; do echo "$f: $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['e2e']['value'])" 2>&1 | tail -1)"; done
bash tools/gpu_ncu_blocks.sh $TAG

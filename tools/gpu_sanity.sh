#!/bin/bash
# Short end-of-session visit: full GPU test-suite, the fused-backward subset with the 512-thread variant off, two bench lines.
TAG=${1:-s1}
O=gpurun_out/$TAG
mkdir -p $O
(timeout 300 python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -60) > $O/pytest_gpu.log
tail -2 $O/pytest_gpu.log
(PCB_BWD_NT512=0 timeout 120 python -m pytest tests/test_mednext_bwd_gpu.py -m gpu -q --maxfail=12 2>&1 | tail -15) > $O/pytest_nt256.log
tail -1 $O/pytest_nt256.log
(timeout 200 python bench.py) > $O/bench_n1.json 2> $O/bench_n1.err
(PCB_BWD_NT512=0 timeout 150 python bench.py --steps 5 --no-cpu-baseline) > $O/bench_nt256.json 2> $O/bench_nt256.err
for f in $O/bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['e2e']['value'])" 2>&1 | tail -1)"; done

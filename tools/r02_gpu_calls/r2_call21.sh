#!/bin/bash
# round 2, call 21: tiled stride-2 depthwise weight gradient + 3-stage level-1 backward at batch 4 — parity, smoke(), per-op times, bench
O=gpurun_out/r2c21
mkdir -p $O
(timeout 600 python -X faulthandler -m pytest tests/test_mednext_bwd_gpu.py tests/test_upstream_wheels_gpu.py -m gpu -q -x --durations=3 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3)
(PCB_BWD_OVERLAP=0 timeout 600 python bench.py --config c2 --profile-ops --no-graph --steps 3 --warmup 3 --no-cpu-baseline --no-e2e) > $O/bench_ops.json 2> $O/bench_ops.err
grep -E "dw_wgrad|mlp_bwd_fused" $O/bench_ops.err | head -14
(timeout 600 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline) > $O/bench_c2.json 2> $O/bench_c2.err
python -c "
import json; d=json.load(open('$O/bench_c2.json')); print('c2', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'roof', d['roofline']['kernel'][:40], d['roofline']['frac'])" 2>&1 | tail -1

#!/bin/bash
# round 2, call 22: tiled transposed stride-2 stencil — parity (forward, stats, both fused-add modes), smoke, ops table, train + infer bench
O=gpurun_out/r2c22
mkdir -p $O
(timeout 600 python -X faulthandler -m pytest tests/test_mednext_gpu.py tests/test_mednext_bwd_gpu.py tests/test_optim_gpu.py -m gpu -q -x --durations=3 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3)
(PCB_BWD_OVERLAP=0 timeout 600 python bench.py --config c2 --profile-ops --no-graph --steps 3 --warmup 3 --no-cpu-baseline --no-e2e) > $O/bench_ops.json 2> $O/bench_ops.err
grep -E "dwconv_fwd:m2|dw_bwd_data:m1|sum of" $O/bench_ops.err | head -10
(timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline) > $O/bench_default.json 2> $O/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c22/bench_default.json"))
print("train", round(d["value"], 2), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2))
i = d["infer"]
print("infer", round(i["value"], 1), round(i["ms_per_step"], 1), "e2e", i["e2e"]["value"])
PY

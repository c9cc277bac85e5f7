#!/bin/bash
# round 2, call 2: whole GPU suite (new parity bounds, new kernels) + the default bench line (train + infer records)
O=gpurun_out/r2c2
mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -q --durations=12 -s 2>&1 | grep -v "^$" | tail -400) > $O/pytest_gpu.log
grep -E "passed|failed|error" $O/pytest_gpu.log | tail -5
grep -E "^FAILED|^ERROR" $O/pytest_gpu.log | head -40
(timeout 600 python bench.py --steps 5 --warmup 3) > $O/bench_default.json 2> $O/bench_default.err
tail -c 600 $O/bench_default.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c2/bench_default.json"))
    print("train", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline", {}).get("value"))
    i = d["infer"]
    print("infer", i["value"], i["ms_per_step"], "e2e", i["e2e"], "launches", i["gpu_launches"], "cpu", i.get("cpu_baseline", {}).get("value"))
    print("roof", d["roofline"]["kernel"], d["roofline"]["frac"], "| infer roof", i["roofline"]["kernel"], i["roofline"]["frac"])
except Exception as e:
    print("bench parse failed", e)
PY

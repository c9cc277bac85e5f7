#!/bin/bash
# round 2, call 27: the default bench line of the final tree (N = 1)
O=gpurun_out/r2c27
mkdir -p $O
(timeout 240 python bench.py --steps 10 --warmup 3) > $O/bench_default.json 2> $O/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c27/bench_default.json"))
print("train", round(d["value"], 2), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "cpu", d["cpu_baseline"]["value"])
i = d["infer"]
print("infer", round(i["value"], 1), round(i["ms_per_step"], 1), "e2e", i["e2e"]["value"], i["execution"].get("remeasured"), "module", i["execution"]["module_path_ms_per_step"])
PY

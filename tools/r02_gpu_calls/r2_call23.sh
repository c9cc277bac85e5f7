#!/bin/bash
# round 2, call 23: what the driver runs at round end — the whole GPU suite, smoke(), the default bench line (+ the 480^3 record)
O=gpurun_out/r2c23
mkdir -p $O
(timeout 1200 python -X faulthandler -m pytest tests -m gpu -q -x --durations=5 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3)
(PCB_BWD_OVERLAP=0 timeout 600 python bench.py --config c2 --profile-ops --no-graph --steps 3 --warmup 3 --no-cpu-baseline --no-e2e) > $O/bench_ops.json 2> $O/bench_ops.err
grep -E "ms/step" $O/bench_ops.err > $O/train_ops.txt; grep -E "dwconv_fwd:m2C64|sum of" $O/train_ops.txt
(timeout 900 python bench.py --steps 10 --warmup 3) > $O/bench_default.json 2> $O/bench_default.err
(timeout 300 python bench.py --mode infer --volume 480 --sw-batch 2 --steps 3 --no-cpu-baseline) > $O/bench_infer480.json 2> $O/bench_infer480.err
(timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > $O/bench_reference.json 2> $O/bench_reference.err
python - <<'PY'
import json
O = "gpurun_out/r2c23/"
d = json.load(open(O + "bench_default.json"))
print("train", round(d["value"], 2), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "cpu", d["cpu_baseline"]["value"], "roof", d["roofline"]["frac"], d["step_roofline"]["frac"], "launches", d["gpu_launches"])
i = d["infer"]
print("infer", round(i["value"], 1), round(i["ms_per_step"], 1), "e2e", i["e2e"]["value"], "cpu", i["cpu_baseline"]["value"], "roof", i["roofline"]["frac"], i["step_roofline"]["frac"])
j = json.load(open(O + "bench_infer480.json")); print("infer480", round(j["value"], 1), round(j["ms_per_step"], 1), "e2e", j["e2e"]["value"])
r = json.load(open(O + "bench_reference.json")); print("reference", r["value"], r["unit"], r["cpu_baseline"]["cores"], "infer", r["infer"]["value"])
PY

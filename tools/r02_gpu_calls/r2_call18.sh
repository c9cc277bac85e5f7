#!/bin/bash
# round 2, call 18: new parity cases (H = 96 fused backward, L up_1, ws2 stages), the driver's default bench line, 480^3
# inference, and the BASELINE configs c3 / c4 / c5 at size on one GPU
O=gpurun_out/r2c18
mkdir -p $O
(timeout 600 python -X faulthandler -m pytest tests/test_mednext_bwd_gpu.py -m gpu -q -x --durations=3 -p no:cacheprovider 2>&1) > $O/pytest_bwd.log
tail -4 $O/pytest_bwd.log
(timeout 200 python tools/time_train_step.py --size L --side 224 --top 8 2>&1 | tail -11) | tee $O/time_L_224.log
(timeout 900 python bench.py --steps 10 --warmup 3) > $O/bench_default.json 2> $O/bench_default.err
(timeout 300 python bench.py --mode infer --volume 480 --sw-batch 2 --steps 3 --no-cpu-baseline) > $O/bench_infer480.json 2> $O/bench_infer480.err
for c in c3 c4; do (timeout 900 python bench.py --config $c --steps 5 --warmup 3) > $O/bench_$c.json 2> $O/bench_$c.err; done
(timeout 900 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu-baseline) > $O/bench_c5.json 2> $O/bench_c5.err
python - <<'PY'
import json
O = "gpurun_out/r2c18/"
def show(f):
    try:
        d = json.load(open(O + f))
    except Exception as e:
        print(f, "FAILED", e); return
    r = d.get("roofline") or {}
    print(f, "|", d["config"]["workload"][:70], "|", round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms | e2e", (d.get("e2e") or {}).get("value"),
          "| cpu", (d.get("cpu_baseline") or {}).get("value"), "| roof", (r.get("kernel") or "")[:36], r.get("frac"), "| step", (d.get("step_roofline") or {}).get("frac"), "| launches", d.get("gpu_launches"))
    if "infer" in d:
        i = d["infer"]; r = i["roofline"]
        print("   infer |", i["config"]["workload"][:70], "|", round(i["value"], 1), i["unit"], round(i["ms_per_step"], 1), "ms | e2e", i["e2e"]["value"], "| cpu", i["cpu_baseline"]["value"],
              "| roof", r["kernel"][:36], r["frac"], "| step", i["step_roofline"]["frac"], "| launches", i["gpu_launches"])
for f in ("bench_default.json", "bench_infer480.json", "bench_c3.json", "bench_c4.json", "bench_c5.json"):
    show(f)
PY

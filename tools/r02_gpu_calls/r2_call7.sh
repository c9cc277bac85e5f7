#!/bin/bash
# round 2, call 7: pipelined epilogues (fwd mlp_fused, level-0 bwd ws): parity suite, A/B timings, default bench
O=gpurun_out/r2c7
mkdir -p $O
(timeout 1200 python -X faulthandler -m pytest tests -m gpu -q -x --durations=5 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
(timeout 200 python tools/time_bwd_ws.py --batch 2 --env PCB_FWD_NOPIPE --modes 1,0 --op mlp_fwd 2>&1 | tail -30) > $O/time_fwd.log
cat $O/time_fwd.log
(timeout 200 python tools/time_bwd_ws.py --batch 2 --env PCB_FWD_LD16 --modes 1 --op mlp_fwd 2>&1 | tail -30) > $O/time_fwd_ld16.log
cat $O/time_fwd_ld16.log
(timeout 300 python tools/time_bwd_ws.py --batch 2 --modes 0,1,2 2>&1 | tail -30) > $O/time_bwd.log
cat $O/time_bwd.log
(timeout 900 python bench.py --steps 5 --warmup 3) > $O/bench_default.json 2> $O/bench_default.err
tail -c 300 $O/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c7/bench_default.json"))
print("train", round(d["value"], 2), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"], "roof", d["roofline"]["kernel"][:40], d["roofline"]["frac"])
i = d["infer"]
print("infer", round(i["value"], 1), round(i["ms_per_step"], 1), "e2e", i["e2e"]["value"], "roof", i["roofline"]["kernel"][:40], i["roofline"]["frac"], i["step_roofline"]["frac"])
PY

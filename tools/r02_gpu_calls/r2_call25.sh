#!/bin/bash
# round 2, call 25 (4 GPUs): the driver's scaling command at N = 4 — interior ranks exchange overlap planes with TWO neighbours
O=gpurun_out/r2c25
mkdir -p $O
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3) > $O/bench_n4.json 2> $O/bench_n4.err
tail -c 300 $O/bench_n4.err
python - <<'PY'
import json
txt = open("gpurun_out/r2c25/bench_n4.json").read()
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print("N=4 train", round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms e2e", d["e2e"]["value"], d["config"]["parallelism"])
i = d["infer"]
print("N=4 infer", round(i["value"], 1), i["unit"], round(i["ms_per_step"], 1), "ms e2e", i["e2e"]["value"], i["config"]["volume"], i["execution"].get("remeasured"))
PY

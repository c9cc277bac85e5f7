#!/bin/bash
# round 2, call 26 (8 GPUs): the driver's scaling command at N = 8
O=gpurun_out/r2c26
mkdir -p $O
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3) > $O/bench_n8.json 2> $O/bench_n8.err
tail -c 300 $O/bench_n8.err
python - <<'PY'
import json
txt = open("gpurun_out/r2c26/bench_n8.json").read()
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print("N=8 train", round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms e2e", d["e2e"]["value"], d["config"]["parallelism"])
i = d["infer"]
print("N=8 infer", round(i["value"], 1), i["unit"], round(i["ms_per_step"], 1), "ms e2e", i["e2e"]["value"], i["config"]["volume"], i["execution"].get("remeasured"))
PY

#!/bin/bash
# round 2, call 19 (2 GPUs): the driver's scaling command at N = 2 (train + sharded inference over NCCL), the 2-rank NCCL parity test
O=gpurun_out/r2c19
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv
(timeout 600 python -m pytest tests/test_sharded_window.py tests/test_lazy_chunked_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4)
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3) > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 400 $O/bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c19/bench_n2.json"))
print("N=2 train", round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms e2e", d["e2e"]["value"], d["config"]["parallelism"], d["execution"]["timed_region"])
i = d["infer"]
print("N=2 infer", round(i["value"], 1), i["unit"], round(i["ms_per_step"], 1), "ms e2e", i["e2e"], i["config"]["parallelism"], i["config"]["volume"])
PY

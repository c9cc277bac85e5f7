#!/bin/bash
# round 2, call 8: per-group MMA issuer warps (fwd: 2, level-0 bwd: 3) — parity, A/B timing, source-level ncu capture
O=gpurun_out/r2c8
mkdir -p $O
(timeout 600 python -X faulthandler -m pytest tests/test_mednext_gpu.py tests/test_mednext_bwd_gpu.py tests/test_optim_gpu.py tests/test_native_gpu.py -m gpu -q --durations=3 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
(timeout 200 python tools/time_bwd_ws.py --batch 2 --env PCB_FWD_NOPIPE --modes 1,0 --op mlp_fwd 2>&1 | tail -30) > $O/time_fwd.log
cat $O/time_fwd.log
(timeout 300 python tools/time_bwd_ws.py --batch 2 --modes 1 2>&1 | tail -30) > $O/time_bwd.log
cat $O/time_bwd.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'mlp_fused_kernel|mlp_bwd_ws_kernel' -c 5 \
  -o $O/mlp_kernels python tools/profile_blocks.py --batch 2 > $O/ncu.log 2>&1
ls -la $O | tail -5
(timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline) > $O/bench_default.json 2> $O/bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c8/bench_default.json"))
print("train", round(d["value"], 2), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"], "roof", d["roofline"]["kernel"][:40], d["roofline"]["frac"])
i = d["infer"]
print("infer", round(i["value"], 1), round(i["ms_per_step"], 1), "e2e", i["e2e"]["value"], "roof", i["roofline"]["kernel"][:40], i["roofline"]["frac"], i["step_roofline"]["frac"])
PY

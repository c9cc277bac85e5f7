#!/bin/bash
# round 2, call 4: re-run the files that failed, then the default bench (native tile loop, fused optimizer) + A/B
O=gpurun_out/r2c4
mkdir -p $O
for f in tests/test_native_gpu.py tests/test_optim_gpu.py tests/test_mednext_gpu.py tests/test_mednext_bwd_gpu.py tests/test_sw_gpu.py tests/test_sharded_window.py; do
  n=$(basename $f .py)
  (timeout 900 python -X faulthandler -m pytest $f -m gpu -q -s --durations=5 -p no:cacheprovider 2>&1) > $O/$n.log
  echo "== $n: $(grep -E '[0-9]+ (passed|failed)|error|Fatal|Segmentation' $O/$n.log | tail -2 | tr '\n' ' ')"
  grep -E "^FAILED|^ERROR" $O/$n.log | head -12
done
(timeout 900 python bench.py --steps 5 --warmup 3) > $O/bench_default.json 2> $O/bench_default.err
tail -c 400 $O/bench_default.err
(timeout 300 env PCB_TORCH_ADAMW=1 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline) > $O/bench_torch_adamw.json 2> $O/bench_torch_adamw.err
python - <<'PY'
import json
for f in ("bench_default", "bench_torch_adamw"):
    try:
        d = json.load(open(f"gpurun_out/r2c4/{f}.json"))
        print(f, "train", round(d["value"], 2), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), d["execution"]["timed_region"], "launches", d["gpu_launches"])
        if "infer" in d:
            i = d["infer"]
            print("   infer", round(i["value"], 1), round(i["ms_per_step"], 1), "e2e", i["e2e"], "launches", i["gpu_launches"], "module-path ms", i["execution"].get("module_path_ms_per_step"))
            print("   infer roof", i["roofline"]["kernel"][:50], i["roofline"]["frac"], i["step_roofline"]["frac"])
    except Exception as e:
        print(f, "parse failed", e)
PY

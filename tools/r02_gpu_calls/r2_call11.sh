#!/bin/bash
# round 2, call 11: warp-specialised cp.async tn_gemm — parity (MedNeXt backward, MONAI U-Net conv taps), A/B per-op times, bench
O=gpurun_out/r2c11
mkdir -p $O
(timeout 600 python -X faulthandler -m pytest tests/test_mednext_bwd_gpu.py tests/test_monai_unet_gpu.py tests/test_optim_gpu.py -m gpu -q -x --durations=3 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
for v in 0 1; do
  echo "== PCB_TN_WS=$v (main stream only)"
  (PCB_TN_WS=$v PCB_BWD_OVERLAP=0 timeout 200 python tools/profile_deep.py --time 2>&1 | grep -E "tn_gemm|gn_bwd|mlp_bwd|mlp_fwd_deep|dw_wgrad") | tee $O/time_deep_ws$v.log
done
(timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline) > $O/bench_c2.json 2> $O/bench_c2.err
(PCB_TN_WS=0 timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline) > $O/bench_c2_old.json 2> $O/bench_c2_old.err
python - <<'PY'
import json
for f in ("bench_c2", "bench_c2_old"):
    d = json.load(open(f"gpurun_out/r2c11/{f}.json"))
    print(f, "train", round(d["value"], 2), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"])
PY

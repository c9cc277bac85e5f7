#!/bin/bash
# round 2, call 14: ws2 operand stages (NB = 1 shapes), LD16, tn_gemm_ws with 8 loader warps; MedNeXt-L sizing
O=gpurun_out/r2c14
mkdir -p $O
(timeout 600 python -X faulthandler -m pytest tests/test_mednext_bwd_gpu.py tests/test_monai_unet_gpu.py -m gpu -q -x --durations=3 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
for v in "PCB_BWD_NST=2" "PCB_BWD_NST=4" "PCB_BWD_NST=4 PCB_BWD_LD16=1" "PCB_BWD_NST=4 PCB_BWD_SPLIT=0"; do
  (env $v timeout 300 python tools/time_bwd_ws.py --batch 2 --modes 2 2>&1 | tail -4 | grep -v "same C32" | sed "s/^/$v /") | tee -a $O/time_bwd.log
done
(PCB_BWD_OVERLAP=0 timeout 200 python tools/profile_deep.py --time 2>&1 | grep -E "tn_gemm") | tee $O/time_deep.log
(timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline) > $O/bench_c2.json 2> $O/bench_c2.err
python -c "
import json; d=json.load(open('$O/bench_c2.json')); print('c2', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'roof', d['roofline']['kernel'][:40], d['roofline']['frac'])" 2>&1 | tail -1
for side in 96 160; do
  (timeout 240 python tools/time_train_step.py --size L --side $side --top 12 2>&1 | tail -18) | tee $O/time_L_$side.log
done

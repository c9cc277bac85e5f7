#!/bin/bash
# round 2, call 1: run every opt-in parity test, A/B the warp-specialised backward kernels and the forward loader depth.
O=gpurun_out/r2c1
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/smi.txt
(PCB_TEST_OPTIN=1 timeout 900 python -m pytest tests/test_mednext_bwd_gpu.py tests/test_mednext_gpu.py -m gpu -q -x -k "warp_specialised or many_tiles or training_step_64 or loader_depth" 2>&1 | tail -40) > $O/pytest_optin.log
tail -5 $O/pytest_optin.log
(timeout 300 python tools/time_bwd_ws.py --batch 2 2>&1 | tail -30) > $O/time_bwd.log
cat $O/time_bwd.log
(timeout 200 python tools/time_bwd_ws.py --batch 2 --env PCB_FWD_LD16 --modes 0,1 --op mlp_fwd 2>&1 | tail -30) > $O/time_fwd_ld16.log
cat $O/time_fwd_ld16.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for v in PCB_BWD_WS=0 PCB_BWD_WS=2 PCB_FWD_LD16=1; do
  (timeout 200 env $v $B) > $O/bench_$v.json 2> $O/bench_$v.err
  echo "$v: $(python -c "import json,sys; d=json.load(open('$O/bench_$v.json')); print(d['value'], d['ms_per_step'], d['execution'])" 2>&1 | tail -1)"
done

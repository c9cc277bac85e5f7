#!/bin/bash
# round 2, call 16: name the MedNeXt-L launch that hangs at 96^3 and bisect it (resident weights / deep path / cp.async staging)
O=gpurun_out/r2c16
mkdir -p $O
for v in "PCB_GW_ASYNC=0" "PCB_NO_BRES=1 PCB_GW_ASYNC=0" "PCB_NO_DEEP=1"; do
  echo "=== $v"
  (env $v PCB_TRACE_OPS=1 PCB_DEBUG_HANG=40 timeout 80 python tools/time_train_step.py --size L --side 96 --top 6 2>&1 | grep -E "^op |^iter|Timeout|sum of" | tail -6) | tee "$O/hang_$(echo $v | tr '= ' '__').log"
done
(timeout 300 python -m pytest tests/test_mednext_gpu.py tests/test_mednext_bwd_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3)
(PCB_BWD_OVERLAP=0 timeout 200 python tools/profile_deep.py --time 2>&1 | grep -E "mlp_bwd|mlp_fwd_deep") | tee $O/time_deep_async.log
(PCB_GW_ASYNC=0 PCB_BWD_OVERLAP=0 timeout 200 python tools/profile_deep.py --time 2>&1 | grep -E "mlp_bwd|mlp_fwd_deep") | tee $O/time_deep_regs.log

#!/bin/bash
# round 2, call 3: GPU suite file by file (a crash in one file cannot hide the others), full logs
O=gpurun_out/r2c3
mkdir -p $O
for f in tests/test_lazy_chunked_gpu.py tests/test_sw_gpu.py tests/test_sharded_window.py tests/test_native_gpu.py tests/test_optim_gpu.py tests/test_tta_gpu.py tests/test_monai_unet_gpu.py tests/test_mednext_gpu.py tests/test_mednext_bwd_gpu.py; do
  n=$(basename $f .py)
  (timeout 900 python -X faulthandler -m pytest $f -m gpu -q -s --durations=5 -p no:cacheprovider 2>&1) > $O/$n.log
  echo "== $n: $(grep -E '[0-9]+ (passed|failed)|error|Fatal|Segmentation' $O/$n.log | tail -2 | tr '\n' ' ')"
  grep -E "^FAILED|^ERROR" $O/$n.log | head -12
done

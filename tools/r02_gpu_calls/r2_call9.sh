#!/bin/bash
# round 2, call 9: whole GPU suite (new: affinity-aware / patch-first TTA)
O=gpurun_out/r2c9
mkdir -p $O
(timeout 1200 python -X faulthandler -m pytest tests -m gpu -q --durations=5 -p no:cacheprovider 2>&1) > $O/pytest_gpu.log
tail -30 $O/pytest_gpu.log

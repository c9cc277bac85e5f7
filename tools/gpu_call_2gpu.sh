#!/bin/bash
# 2-GPU visit: data-parallel train bench (two graphs around the eager NCCL all-reduce), eager DDP for comparison,
# z-slab sharded inference bench.  Output: gpurun_out/<tag>/
TAG=${1:-g2}
O=gpurun_out/$TAG
mkdir -p $O
T="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
($T bench.py --gpus 2 --steps 5 --warmup 3) > $O/train_n2.json 2> $O/train_n2.err
($T bench.py --gpus 2 --steps 5 --warmup 3 --no-graph) > $O/train_n2_eager.json 2> $O/train_n2_eager.err
(PCB_GRAPH_DDP=1 $T bench.py --gpus 2 --steps 5 --warmup 3) > $O/train_n2_onegraph.json 2> $O/train_n2_onegraph.err
($T bench.py --gpus 2 --mode infer --steps 2 --warmup 3) > $O/infer_n2.json 2> $O/infer_n2.err
($T bench.py --gpus 2 --impl reference --steps 1 --warmup 0) > $O/ref_n2.json 2> $O/ref_n2.err
for f in $O/*.json; do echo "$f: $(python -c "import json,sys; d=json.load(open('$f')); print(d.get('value'), d.get('ms_per_step'), d.get('execution'), d.get('e2e'))" 2>&1 | tail -1)"; done
tail -3 $O/*.err

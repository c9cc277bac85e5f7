"""Run the deep part of one MedNeXt-S training step (level-2 block, up_1 with its skip; forward + backward) so that ncu can
capture the deep launch classes (gemm_ws_kernel, tn_gemm_kernel) in one short pass:

    ncu --set full --clock-control none --import-source on -k regex:tn_gemm -c 12 -o gpurun_out/deep python tools/profile_deep.py

Usage: python tools/profile_deep.py [--batch 4] [--side 40] [--time]   (--time: CUDA-event times of the weight-gradient GEMMs)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_connectomics_b200 import _lib as L  # noqa: E402
from pytorch_connectomics_b200.architectures import mednext as PM  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--side", type=int, default=40)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--time", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = PM.MedNeXtBlock(128, 128, 2, 3).to(dev).train()
up = PM.MedNeXtUpBlock(128, 64, 2, 3, do_res=True).to(dev).train()


def cl(n, s, c):
    t = torch.randn(n, s, s, s, c, device=dev).bfloat16()
    t._pcb_cl = True
    return t.requires_grad_(True)


x, skip = cl(a.batch, a.side, 128), cl(a.batch, 2 * a.side, 64)
for it in range(a.iters + (2 if a.time else 0)):
    if a.time and it == 2:
        L.prof_start([])
    y = up(blk(x), skip)
    y.backward(torch.ones_like(y))
if a.time:
    for k, v in sorted(L.prof_stop().items()):
        v = sorted(v)
        print(f"{k:60s} n={len(v):3d} median {v[len(v) // 2]:.3f} ms  sum {sum(v):.3f} ms")
torch.cuda.synchronize()
print("done")

"""A/B the level-0 fused MLP backward: default 256-thread kernel vs the opt-in warp-specialised one (PCB_BWD_WS=1).
CUDA-event time of the op alone (pytorch_connectomics_b200._lib.prof hooks), 160^3 x 32 channels, batch --batch."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_connectomics_b200 import _lib as L  # noqa: E402
from pytorch_connectomics_b200.architectures import mednext as PM  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--side", type=int, default=160)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = PM.MedNeXtBlock(32, 32, 2, 3).to(dev).train()
x = torch.randn(a.batch, a.side, a.side, a.side, 32, device=dev).bfloat16()
x._pcb_cl = True
x.requires_grad_(True)
g = None
for mode in ("0", "1", "0", "1"):
    os.environ["PCB_BWD_WS"] = mode
    out = blk(x)
    if g is None:
        g = torch.randn_like(out)
    out.backward(g)                     # warm-up
    L.prof_start([])
    for _ in range(a.iters):
        out = blk(x)
        out.backward(g)
    t = L.prof_stop()
    k = [k for k in t if k.startswith("mlp_bwd_fused")][0]
    ts = sorted(t[k])
    print(f"PCB_BWD_WS={mode}: {k} median {ts[len(ts) // 2]:.3f} ms  min {ts[0]:.3f} ms  (batch {a.batch})")

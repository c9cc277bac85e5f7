"""A/B the fused MLP backward: default kernel vs the warp-specialised ones (PCB_BWD_WS=1/2) on the level-0, level-1,
down_0 and up_0 launch classes.  CUDA-event time of the op alone (pytorch_connectomics_b200._lib.prof hooks)."""
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
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--modes", default="0,1,2")
ap.add_argument("--env", default="PCB_BWD_WS")
ap.add_argument("--op", default="mlp_bwd_fused")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
cases = [("same", 32, 32, a.side), ("same", 64, 64, a.side // 2), ("down", 32, 64, a.side), ("up", 64, 32, a.side // 2)]
for kind, cin, cout, side in cases:
    cls = {"same": PM.MedNeXtBlock, "down": PM.MedNeXtDownBlock, "up": PM.MedNeXtUpBlock}[kind]
    blk = cls(cin, cout, 2, 3).to(dev).train()
    x = torch.randn(a.batch, side, side, side, cin, device=dev).bfloat16()
    x._pcb_cl = True
    x.requires_grad_(True)
    g = None
    for mode in a.modes.split(","):
        os.environ[a.env] = mode
        try:
            out = blk(x)
            if g is None:
                g = torch.randn_like(out)
            out.backward(g)                     # warm-up
            L.prof_start([])
            for _ in range(a.iters):
                out = blk(x)
                out.backward(g)
            t = L.prof_stop()
            for k in sorted(t):
                if k.startswith(a.op):
                    ts = sorted(t[k])
                    print(f"{kind} C{cin}->{cout} {a.env}={mode}: {k} median {ts[len(ts) // 2]:.3f} ms  min {ts[0]:.3f} ms  (batch {a.batch})", flush=True)
        except Exception as exc:
            print(f"{kind} C{cin}->{cout} {a.env}={mode}: FAILED {exc!r}", flush=True)
    del blk, x, g
    torch.cuda.empty_cache()

"""Time the MedNeXt forward (inference) on one GPU: whole-net CUDA-event time + optional per-op split.
Usage: python tools/time_fwd.py [--size S] [--side 160] [--batch 1] [--iters 10] [--split]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_connectomics_b200.architectures import mednext as PM  # noqa: E402
from pytorch_connectomics_b200.architectures import _mednext_ops as ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="S")
ap.add_argument("--side", type=int, default=160)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--split", action="store_true")
a = ap.parse_args()

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = PM.create_mednext_v1(1, 1, a.size, 3, False).eval().to(dev)
x = torch.rand(a.batch, 1, a.side, a.side, a.side, device=dev).half()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

with torch.no_grad():
    for _ in range(a.warmup):
        net(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        net(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"MedNeXt-{a.size} fwd {a.batch}x{a.side}^3: median {med:.3f} ms  min {ts[0]:.3f} ms  "
          f"=> {a.batch * 1000.0 / med:.1f} sub-vol/s, {a.batch * a.side ** 3 / med / 1e3:.1f} Mvox/s")

    if a.split:
        # per-op split using events around the raw launches of the first encoder block at full res
        import ctypes
        from pytorch_connectomics_b200 import _lib as L
        f = ops.stem_forward(x, net.stem.weight, net.stem.bias)
        blk = net.enc_block_0[0]
        def t(fn, n=10):
            fn(); torch.cuda.synchronize()
            r = []
            for _ in range(n):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                r.append(e0.elapsed_time(e1))
            r.sort(); return r[len(r) // 2]
        print("stem            %.3f ms" % t(lambda: ops.stem_forward(x, net.stem.weight, net.stem.bias)))
        print("level-0 block   %.3f ms" % t(lambda: blk(f)))
        d = net.down_0(f)
        print("down_0          %.3f ms" % t(lambda: net.down_0(f)))
        print("level-1 block   %.3f ms" % t(lambda: net.enc_block_1[0](d)))
        print("up_0 (+skip)    %.3f ms" % t(lambda: net.up_0(d, f)))
        print("head            %.3f ms" % t(lambda: net.out_0(f, torch.float16)))

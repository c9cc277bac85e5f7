"""Run the full-resolution part of one MedNeXt-S training step (stem, one level-0 block, down_0, one level-1 block,
up_0, head; forward + backward) so that `ncu --set full` can capture every level-0/1 launch class in one short pass:

    ncu --set full --clock-control none --import-source on -k regex:pcb -o gpurun_out/blocks python tools/profile_blocks.py

Usage: python tools/profile_blocks.py [--side 160] [--batch 1] [--iters 1] [--time]
With --time the script prints CUDA-event times per section instead (do not combine with ncu)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_connectomics_b200.architectures import mednext as PM  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=160)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--time", action="store_true")
a = ap.parse_args()

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = PM.create_mednext_v1(1, 1, "S", 3, False).train().to(dev)
x = torch.rand(a.batch, 1, a.side, a.side, a.side, device=dev).half()


def section(name, fn):
    if not a.time:
        return fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1):8.3f} ms")
    return r


for it in range(a.iters + (1 if a.time else 0)):
    if a.time:
        print(f"--- iteration {it}{' (warm-up)' if it == 0 else ''}")
    f = section("stem", lambda: PM.ops.stem_apply(x, net.stem.weight, net.stem.bias))
    r0 = section("enc_block_0[0] fwd", lambda: net.enc_block_0[0](f))
    d = section("down_0 fwd", lambda: net.down_0(r0))
    r1 = section("enc_block_1[0] fwd", lambda: net.enc_block_1[0](d))
    u = section("up_0 fwd (+skip)", lambda: net.up_0(r1, r0))
    o = section("out_0 fwd", lambda: net.out_0(u, torch.float16))
    g = torch.ones_like(o)
    # backward section by section (autograd graph is a chain; time it as a whole, the per-op split comes from ncu)
    section("backward (all of the above)", lambda: o.backward(g))
    net.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("done")

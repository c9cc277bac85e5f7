"""One eager training step of a MedNeXt (fwd + BCE + bwd) at a given size/crop with per-op CUDA-event times — used to size the
MedNeXt-L config (BASELINE configs[3]) and to find its slow launch classes.
Usage: python tools/time_train_step.py --size L --side 224 [--batch 1] [--top 25]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_connectomics_b200 import _lib as L  # noqa: E402
from pytorch_connectomics_b200.architectures import mednext as PM  # noqa: E402

if os.environ.get("PCB_DEBUG_HANG"):      # find a hung launch: blocking launches + a Python traceback after N seconds
    import faulthandler
    os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
    faulthandler.dump_traceback_later(int(os.environ["PCB_DEBUG_HANG"]), exit=True)

if os.environ.get("PCB_TRACE_OPS"):       # print every op scope as it starts (the last line names a hung launch)
    _enter = L.prof.__enter__

    def _traced(self):
        print("op", self.name, flush=True)
        return _enter(self)

    L.prof.__enter__ = _traced

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="L")
ap.add_argument("--side", type=int, default=224)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--top", type=int, default=25)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = PM.create_mednext_v1(1, 1, a.size, 3, False).train().to(dev)
x = torch.rand(a.batch, 1, a.side, a.side, a.side, device=dev).half()
t = (torch.rand(a.batch, 1, a.side, a.side, a.side, device=dev) > 0.85).float()
bce = torch.nn.functional.binary_cross_entropy_with_logits
for it in range(2):
    if it == 1:
        L.prof_start([])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = net(x)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    bce(out.float(), t).backward()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"iter {it}: forward {1e3 * (t1 - t0):.1f} ms, loss+backward {1e3 * (t2 - t1):.1f} ms, peak mem "
          f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    net.zero_grad(set_to_none=True)
prof = L.prof_stop()
rows = sorted(((sum(v), len(v), k) for k, v in prof.items()), reverse=True)
print(f"sum of timed ops {sum(r[0] for r in rows):.1f} ms")
for s, n, k in rows[:a.top]:
    print(f"{k:48s} n={n:3d} {s:9.3f} ms")

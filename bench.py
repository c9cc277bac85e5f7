#!/usr/bin/env python
"""pcb200 benchmark — prints ONE JSON line (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (training step, config 2)
    python bench.py --impl reference --steps K --warmup W    # reference CPU arm (oracle port)
    python bench.py --mode infer ...                         # sliding-window inference workload

Workload (BASELINE.json configs[1]): MedNeXt-S, 1-channel 160^3 crops, bf16 compute, per-GPU batch 4 (--batch),
BCE-with-logits + Dice loss, AdamW(lr 1e-3, wd 0.01) — one "step" = forward + loss + backward + gradient
all-reduce (N>1) + optimizer step on one synthetic batch per GPU.  `value` = sub-volumes/s over all
ranks with inputs resident in HBM; `e2e` = the same step through the public API with the batch copied from
pinned host memory every step and the loss read back.  Inputs are regenerated (different tensors) each step
and each step touches >1 GB of activations, far beyond the 126 MB L2, so no explicit L2 flush is needed.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace as NS

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SIDE = 160
METRIC_TRAIN = "training sub-volumes/sec"
METRIC_INFER = "inference Mvoxels/sec"


def cfg_mednext(size="S", out_channels=1):
    return NS(model=NS(arch=NS(type="mednext"), in_channels=1, out_channels=out_channels,
                       mednext=NS(size=size, kernel_size=3, checkpoint_style="outside_block"),
                       loss=NS(deep_supervision=False)))


def bce_dice_loss(logits, target):
    """WeightedBCE + Dice as in tutorials/mito_lucchi++ (loss math stays PyTorch — SURVEY §2)."""
    logits = logits.float()
    bce = torch.nn.functional.binary_cross_entropy_with_logits(logits, target)
    p = torch.sigmoid(logits)
    inter = (p * target).sum()
    dice = 1.0 - (2.0 * inter + 1.0) / (p.sum() + target.sum() + 1.0)
    return bce + dice


# ----------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ----------------------------------------------------------------------------- reference / CPU arm
def cpu_train_rate(steps: int, warmup: int, side: int = 64):
    """The reference's own path on host cores: the oracle MedNeXt-S (pure-torch restatement of the
    un-vendored nnunet_mednext modules the reference builds) — fp32 forward + loss + backward + AdamW
    on a bounded sample (one `side`^3 crop per step), scaled to 160^3 sub-volumes by voxel count."""
    from oracle.mednext_oracle import create_mednext_v1
    cores = min(os.cpu_count() or 1, 32)   # oneDNN stops scaling (and regresses) beyond ~32 threads at this size
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = create_mednext_v1(1, 1, "S", 3, False)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=0.01)
    ts = []
    for i in range(warmup + steps):
        x = torch.rand(1, 1, side, side, side)
        t = (torch.rand(1, 1, side, side, side) > 0.85).float()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = bce_dice_loss(net(x), t)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    sec = sum(ts) / len(ts)
    scale = (side / SIDE) ** 3
    return {"value": scale / sec, "unit": "sub-volumes/s", "cores": cores, "kind": "port",
            "sample": f"oracle MedNeXt-S fp32 train step on one {side}^3 crop per step ({sec:.2f} s/step), "
                      f"scaled by ({side}/{SIDE})^3 to 160^3 sub-volumes", "sec_per_step": sec}


def cpu_infer_rate(steps: int, warmup: int, side: int = 64):
    from oracle.mednext_oracle import create_mednext_v1
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = create_mednext_v1(1, 1, "S", 3, False).eval()
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            x = torch.rand(1, 1, side, side, side)
            t0 = time.perf_counter()
            net(x)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    # 50 % overlap => every output voxel is covered by 8 tiles in the interior
    return {"value": side ** 3 / sec / 8.0 / 1e6, "unit": "Mvox/s", "cores": cores, "kind": "port",
            "sample": f"oracle MedNeXt-S fp32 forward on one {side}^3 tile per step ({sec:.2f} s), /8 tile coverage",
            "sec_per_step": sec}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(a.steps, 3)), max(0, min(a.warmup, 1))
    train = a.mode == "train"
    r = cpu_train_rate(steps, warm) if train else cpu_infer_rate(steps, warm)
    out = {"impl": "reference", "metric": METRIC_TRAIN if train else METRIC_INFER, "value": r["value"],
           "unit": r["unit"], "n_gpus": a.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["sec_per_step"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(a, 1),
           "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "reference third-party nets (nnunet_mednext) are not installable offline; this is the oracle port "
                   "of the reference's PyTorch CPU path on all host threads"}
    print(json.dumps(out))


def workload_config(a, world):
    if a.mode == "train":
        return {"workload": f"MedNeXt-S k3 train step, {a.batch}x1x{SIDE}^3 crops per GPU, BCE+Dice, AdamW (BASELINE configs[1])",
                "global_batch": world * a.batch, "crop": [SIDE] * 3, "parallelism": f"dp{world}",
                "l2": "inputs+activations >> 126 MB L2 per step (no explicit flush)"}
    return {"workload": f"MedNeXt-S sliding-window inference, {a.volume * world}x{a.volume}x{a.volume} volume ({a.volume}^3 per GPU), {SIDE}^3 tiles, 50% overlap, bump",
            "volume": [a.volume * world, a.volume, a.volume], "tile": [SIDE] * 3, "overlap": 0.5,
            "parallelism": "single GPU" if world == 1 else f"one volume, z-slab shards x{world} + neighbour overlap exchange (NCCL send/recv)",
            "l2": "volume+accumulators >> 126 MB L2 (no explicit flush)"}



# ----------------------------------------------------------------------------- roofline leg
# op-name prefix (pytorch_connectomics_b200 profiler hooks) -> kernel symbol, algorithmic bytes per voxel-channel
KERNELS = {
    "mlp_fwd": "pcb::mlp_fused_kernel / pcb::mlp_kernel (GN-apply -> GEMM -> GELU -> GEMM + residual, tcgen05)",
    "mlp_bwd_fused": "pcb::mlp_bwd_fused_kernel (dgrad + both pointwise wgrads in TMEM, tcgen05)",
    "mlp_bwd": "pcb::mlp_bwd_kernel (dgrad, tcgen05)",
    "dwconv_fwd": "pcb::dwconv_same_tiled_kernel<3> / dwconv_kernel (depthwise stencil + GN statistics)",
    "dw_bwd_data": "pcb::dwconv_same_tiled_kernel<3> / dwconv_kernel (stencil data gradient)",
    "dw_wgrad": "pcb::dw_wgrad_same_tiled_kernel<3> / dw_wgrad_kernel (depthwise weight gradient)",
    "tn_gemm": "pcb::tn_gemm_kernel (split-K wgrad GEMM, tcgen05 MN-major)",
    "gn_bwd": "pcb::gn_dy_kernel (GroupNorm backward)",
}
NCU_TRAFFIC = {  # DRAM bytes (read + write) PER SAMPLE of one launch: dram__bytes.sum.per_second x gpu__time_duration from the
    # batch-1 block capture profiles/r01_blocks_final.ncu-rep (raw page: profiles/r01_blocks_final_raw.csv)
    "mlp_bwd_fused:m0C32H64Co32V4096000": 759.3e6,
    "mlp_bwd_fused:m2C64H128Co32V4019679": 1253.9e6,
    "mlp_fwd:m0C32H64Co32V4096000": 758.6e6,
    "mlp_fwd:m2C64H128Co32V4096000": 1084.8e6,
    "dwconv_fwd:m0C32V4096000": 487.3e6,
    "dw_bwd_data:m0C32V4096000": 763.4e6,
    "dw_wgrad:m0C32V4096000": 604.8e6,
    "gn_bwd:C32V4096000": 733.1e6,
}


def _op_bytes(op: str, key: str, batch: int):
    """Algorithmic HBM bytes of one launch from the shape encoded in the profiler key (bf16 activations)."""
    import re
    f = {m[0]: int(m[1]) for m in re.findall(r"([A-Za-z]+?)(\d+)", key.split(":", 1)[1])}
    c, v, co = f.get("C", 0), f.get("V", 0), f.get("Co", f.get("C", 0))
    if op == "mlp_fwd":
        return batch * v * 2 * (c + 2 * co)            # read y, read residual/skip, write out
    if op in ("mlp_bwd_fused", "mlp_bwd"):
        return batch * v * 2 * (2 * c + co)            # read y, read dOut, write dYhat
    if op == "dw_bwd_data" and ":m0" in key:
        return batch * v * 2 * 3 * c                   # read dy + fused residual gradient, write dx
    if op in ("dwconv_fwd", "dw_bwd_data", "dw_wgrad"):
        return batch * v * 2 * 2 * c                   # read + write (or read two tensors)
    if op == "gn_bwd":
        return batch * v * 2 * 3 * c
    return None


def build_roofline(prof, a, peak_gbs, measured, step_ms, nsteps):
    """Dominant kernel = op type with the largest share of the (eager, event-instrumented) step; its roofline
    numbers are quoted on its biggest launch class (level 0) with the live CUDA-event launch durations."""
    if not prof:
        return None
    groups = {}
    for key, dts in prof.items():
        op = key.split(":", 1)[0]
        g = groups.setdefault(op, {"ms": 0.0, "classes": {}})
        g["ms"] += sum(dts) / nsteps
        g["classes"][key] = dts
    ranked = sorted(groups.items(), key=lambda kv: -kv[1]["ms"])
    batch = a.batch if a.mode == "train" else 2
    rows = []
    for op, g in ranked[:4]:
        key, dts = max(g["classes"].items(), key=lambda kv: sum(kv[1]))
        avg_ms = sum(dts) / len(dts)
        nbytes = _op_bytes(op, key, batch)
        ach = nbytes / (avg_ms / 1e3) / 1e9 if nbytes else None
        rows.append({"kernel": KERNELS.get(op, op), "launch_class": key, "bound": "hbm", "achieved": ach, "peak": peak_gbs,
                     "unit": "GB/s", "frac": (ach / peak_gbs) if ach else None, "traffic": (NCU_TRAFFIC[key] * batch) if key in NCU_TRAFFIC else None,
                     "avg_launch_ms": avg_ms, "launches_timed": len(dts), "algorithmic_bytes": nbytes,
                     "op_ms_per_step": g["ms"], "share_of_step": g["ms"] / step_ms})
    roof = dict(rows[0])
    roof["peak_source"] = "MEASURED_PEAKS.json (of measured)" if measured else "fallback 6650 GB/s (of fallback)"
    roof["traffic_source"] = "ncu DRAM bytes (read + write) per launch from the batch-1 capture in profiles/ x batch; null = not captured"
    roof["others"] = rows[1:]
    return roof

def step_roofline(mode: str, units_per_step: float, ms_per_step: float, hbm_gbs: float, bf16_tflops: float):
    """Whole-step roofline (BASELINE.md §2 / SURVEY §8d): MedNeXt-S k3 at 160^3 needs 522 GFLOP and, with fused blocks,
    5.49 GB of HBM traffic per forward; a training step is 3x that.  Time per unit = max(bytes / HBM, FLOPs / tensor peak)
    with the MEASURED peaks; `frac` = roofline time / achieved time of the whole step (not of one kernel)."""
    passes = 3.0 if mode == "train" else 1.0
    t_unit = max(passes * 5.49e9 / (hbm_gbs * 1e9), passes * 522e9 / (bf16_tflops * 1e12))      # s per 160^3 sub-volume / tile
    roof_ms = units_per_step * t_unit * 1e3
    return {"unit": "160^3 sub-volume (train: fwd+bwd = 3 x fwd)" if mode == "train" else "160^3 tile forward",
            "bytes_per_unit": passes * 5.49e9, "flops_per_unit": passes * 522e9, "units_per_step": units_per_step,
            "roofline_ms_per_step": roof_ms, "frac": roof_ms / ms_per_step if ms_per_step > 0 else None,
            "bound": "hbm" if passes * 5.49e9 / (hbm_gbs * 1e9) >= passes * 522e9 / (bf16_tflops * 1e12) else "tensor"}


# ----------------------------------------------------------------------------- our arm
T0 = time.time()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="pcb200", choices=["pcb200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--batch", type=int, default=4, help="sub-volumes per GPU per step (tutorials/mito_lucchi++ trains with 4)")
    ap.add_argument("--volume", type=int, default=480)
    ap.add_argument("--sw-batch", type=int, default=4, help="windows per network call in --mode infer")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-ops", action="store_true", help="time every pcb200 op with CUDA events (stderr table)")
    ap.add_argument("--no-graph", action="store_true", help="run the timed steps eagerly instead of as one CUDA graph")
    a = ap.parse_args()
    a.warmup = max(3, a.warmup) if a.impl == "pcb200" else a.warmup
    if a.impl == "reference":
        return run_reference(a)

    import torch.distributed as dist
    from pytorch_connectomics_b200 import _lib as L

    def dbg(msg):
        if os.environ.get("PCB_BENCH_DEBUG"):
            print(f"[bench r{os.environ.get('RANK', '0')} +{time.time() - T0:7.1f}s] {msg}", file=sys.stderr, flush=True)

    from pytorch_connectomics_b200.architectures import build_model
    from pytorch_connectomics_b200.training import FlatGradArena

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout (one JSON line contract)
        dist.init_process_group("nccl", device_id=dev)
        dbg("process group up")
    if not os.path.exists(L.LIB_PATH):
        L.build()
    if L.lib().pcb_device_ok() != 1:
        raise RuntimeError(L.lib().pcb_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    torch.manual_seed(1234 + rank)
    model = build_model(cfg_mednext("S", 1)).to(dev)
    clocks = Clocks(local) if rank == 0 else None

    if a.mode == "train":
        model.train()
        arena = FlatGradArena(model.parameters())
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01, fused=True, capturable=True)
        nb = a.batch
        shape = (nb, 1, SIDE, SIDE, SIDE)
        # a small pool of distinct resident inputs (fp16 volumes, as the data pipeline delivers them)
        pool = [(torch.rand(shape, device=dev).half(), (torch.rand(shape, device=dev) > 0.85).float()) for _ in range(4)]

        def step(x, t):
            arena.zero()
            loss = bce_dice_loss(model(x), t)
            loss.backward()
            arena.allreduce()
            opt.step()
            return loss

        for i in range(a.warmup):
            step(*pool[i % 4])
            dbg(f"warmup step {i} enqueued")
        barrier()
        dbg("warmup done")
        # (1) instrumented eager pass: per-launch CUDA events around the dominant kernel (roofline leg)
        nprof = a.steps if (a.no_graph or a.profile_ops) else min(3, a.steps)
        L.prof_start([])   # every pcb200 op gets a CUDA-event pair
        l0 = L.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for i in range(nprof):
            step(*pool[i % 4])
        e1.record()
        barrier()
        w1 = time.time()
        launches_per_step = (L.launch_count() - l0) // max(1, nprof)
        prof = L.prof_stop()
        ms_eager = max_over_ranks(e0.elapsed_time(e1)) / nprof
        # (2) timed region: the same step captured once as a CUDA graph and replayed K times
        graphed = None
        run = step
        dbg(f"eager pass done: {ms_eager:.2f} ms/step")
        # data-parallel runs: two graphs around an eagerly launched NCCL all-reduce (PCB_GRAPH_DDP=1: one graph, NCCL captured)
        if not (a.no_graph or a.profile_ops):
            try:
                from pytorch_connectomics_b200.training import GraphedTrainStep
                graphed = GraphedTrainStep(model, bce_dice_loss, opt, arena, pool[0][0], pool[0][1], warmup=1,
                                           split_collective=world > 1 and not os.environ.get("PCB_GRAPH_DDP"))
                run = graphed
            except Exception as exc:   # keep the bench alive; say so in the JSON line
                print(f"[bench] CUDA-graph capture failed, timing eager steps: {exc!r}", file=sys.stderr)
                graphed = None
                arena.rebind()
        for i in range(2):
            run(*pool[i % 4])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for i in range(a.steps):
            run(*pool[i % 4])
        e1.record()
        barrier()
        w1 = time.time()
        launches = launches_per_step * a.steps
        ms = max_over_ranks(e0.elapsed_time(e1))
        value = world * nb * a.steps / (ms / 1e3)
        dbg(f"timed region done: {ms / a.steps:.2f} ms/step")

        # ---- end to end: pinned host batch -> H2D -> step -> loss D2H, every step
        hx = [torch.rand(shape).half().pin_memory() for _ in range(2)]
        ht = [(torch.rand(shape) > 0.85).float().pin_memory() for _ in range(2)]
        h2d = hx[0].numel() * 2 + ht[0].numel() * 4

        def e2e_step(i):
            if graphed is not None:      # pinned host -> static device buffers -> graph replay -> loss D2H
                return float(graphed(hx[i % 2], ht[i % 2]).item())
            x = hx[i % 2].to(dev, non_blocking=True)
            t = ht[i % 2].to(dev, non_blocking=True)
            return float(step(x, t).item())

        for i in range(2):
            e2e_step(i)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(a.steps):
            e2e_step(i)
        f1.record()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))
        dbg("e2e done")
        e2e = {"value": world * nb * a.steps / (ms_e2e / 1e3), "unit": "sub-volumes/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
        metric, unit = METRIC_TRAIN, "sub-volumes/s"
    else:
        model.eval()
        from pytorch_connectomics_b200.inference.window import EagerSlidingWindowEngine
        eng = EagerSlidingWindowEngine(roi_size=(SIDE,) * 3, sw_batch_size=a.sw_batch, overlap=0.5, mode="bump",
                                       padding_mode="constant", cval=0.0)
        net = lambda t: model(t)  # noqa: E731
        if world > 1:
            # ONE volume of world x `volume` planes, z-slab sharded with a neighbour exchange of the overlap planes
            # (inference/sharded.py); every rank holds the same synthetic volume and stages only its slab
            from pytorch_connectomics_b200.inference.sharded import ZSlabShardedEngine
            torch.manual_seed(99)
            vol = torch.rand(1, 1, a.volume * world, a.volume, a.volume, device=dev).half()
            sh = ZSlabShardedEngine(roi_size=(SIDE,) * 3, sw_batch_size=a.sw_batch, overlap=0.5, mode="bump",
                                    padding_mode="constant", cval=0.0, device=dev)

            def step():
                with torch.no_grad():
                    return sh(vol, net)
        else:
            vol = torch.rand(1, 1, a.volume, a.volume, a.volume, device=dev).half()

            def step():
                with torch.no_grad():
                    return eng(inputs=vol, network=net)

        for _ in range(max(1, a.warmup // 3)):
            step()
        barrier()
        L.prof_start([])
        l0 = L.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        barrier()
        w1 = time.time()
        launches = L.launch_count() - l0
        prof = L.prof_stop()
        ms = max_over_ranks(e0.elapsed_time(e1))
        value = world * a.steps * a.volume ** 3 / (ms / 1e3) / 1e6
        if world > 1:
            hv = torch.rand(1, 1, a.volume * world, a.volume, a.volume).half().pin_memory()
        else:
            hv = torch.rand(1, 1, a.volume, a.volume, a.volume).half().pin_memory()
        eng_cpu = EagerSlidingWindowEngine(roi_size=(SIDE,) * 3, sw_batch_size=a.sw_batch, overlap=0.5, mode="bump",
                                           padding_mode="constant", cval=0.0, sw_device=dev, output_device="cpu")
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        with torch.no_grad():
            if world > 1:      # host volume -> this rank's slab H2D -> windows -> exchange -> own planes D2H
                part, own = sh(hv, net)
                out = part.cpu() if part is not None else torch.empty(0)
                h2d_bytes = 0 if part is None else (own[1] - own[0] + SIDE // 2) * a.volume * a.volume * 2
            else:
                out = eng_cpu(inputs=hv, network=net)
                h2d_bytes = hv.numel() * 2
        f1.record()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))
        e2e = {"value": world * a.volume ** 3 / (ms_e2e / 1e3) / 1e6, "unit": "Mvox/s",
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": out.numel() * out.element_size()}
        metric, unit = METRIC_INFER, "Mvox/s"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if a.profile_ops:
        tot = sum(sum(v) for v in prof.values())
        for k, v in sorted(prof.items(), key=lambda kv: -sum(kv[1])):
            print(f"{k:40s} n={len(v) // a.steps:3d}/step  {sum(v) / a.steps:8.3f} ms/step  avg {sum(v) / len(v):7.3f} ms",
                  file=sys.stderr)
        print(f"{'sum of timed ops':40s} {tot / a.steps:8.3f} ms/step of {ms / a.steps:.3f}", file=sys.stderr)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    roof = build_roofline(prof, a, peak_gbs, bool(peaks), ms_eager if a.mode == "train" else ms / a.steps,
                          nprof if a.mode == "train" else a.steps)
    out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic", "config": workload_config(a, world), "e2e": e2e,
           "gpu_launches": int(launches), "roofline": roof,
           "execution": ({"timed_region": ("cuda_graph_replay" if graphed.graph_opt is None else "cuda_graph_replay x2 around eager NCCL all-reduce") if graphed is not None else "eager",
                          "eager_ms_per_step": ms_eager, "roofline_timed_in": f"instrumented eager pass of {nprof} steps"}
                         if a.mode == "train" else {"timed_region": "eager"}),
           "clocks": clocks.window(w0, w1) if clocks else None}
    if clocks:
        clocks.stop()
    try:        # whole-step roofline beside the dominant-kernel one (never allowed to cost the JSON line)
        tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1366.0)))
        if a.mode == "train":
            out["step_roofline"] = step_roofline("train", a.batch, ms / a.steps, peak_gbs, tf)      # per GPU
        elif world == 1:
            ntile = (max(a.volume, SIDE) - SIDE + SIDE // 2 - 1) // (SIDE // 2) + 1
            out["step_roofline"] = step_roofline("infer", ntile ** 3, ms / a.steps, peak_gbs, tf)
    except Exception as exc:
        out["step_roofline"] = {"error": repr(exc)}
    if world == 1 and not a.no_cpu_baseline:
        r = cpu_train_rate(1, 0) if a.mode == "train" else cpu_infer_rate(1, 0)
        out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

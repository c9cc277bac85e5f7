#!/usr/bin/env python
"""pcb200 benchmark — prints ONE JSON line (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W               # our arm: C2 training step + C5-geometry inference record
    python bench.py --config c2|c3|c4|c5 ...                    # one BASELINE.json config on its own
    python bench.py --impl reference --steps K --warmup W       # reference CPU arm (oracle port, real 160^3 crops)

BASELINE.json's metric has two halves — training sub-volumes/s and inference Mvoxels/s — so the default line carries
both: the top-level `metric/value/roofline/e2e/cpu_baseline` are the TRAINING step of configs[1] (MedNeXt-S, 1-channel
160^3 crops, bf16 compute, per-GPU batch 4, BCE+Dice, grad-norm clip 1.0 + AdamW; timed over exactly K steps), and `infer` is a second record
of the same shape (metric, value, unit, ms_per_step, steps, config, roofline, e2e, cpu_baseline) for sliding-window
inference with configs[4]'s geometry (160^3 tiles, 50 % overlap, bump blending) on the largest cubic volume per GPU that
keeps the whole run within a few minutes (--volume, default 640; N GPUs: ONE (N*volume) x volume x volume volume,
z-slab sharded).  `--config c5` runs the full 2048^3 volume; `--config c3` / `c4` the 3-channel and MedNeXt-L training
configs.  `value` = whole-job throughput with inputs resident in HBM; `e2e` = the same work through the public API with
the inputs in pinned host memory and the result read back, copies inside the timed region.  Every step streams far more
than the 126 MB L2 (activations > 1 GB), inputs rotate over a pool of distinct tensors: no explicit L2 flush (said in
`config.l2`).
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

# BASELINE.json configs[1..4] (configs[0] is the CPU plumbing case, a parity test — tests/test_monai_unet_gpu.py)
TRAIN_CONFIGS = {
    "c2": dict(size="S", out_channels=1, crop=160, batch=4,
               name="MedNeXt-S k3 train step, {b}x1x160^3 crops per GPU, BCE+Dice, AdamW (BASELINE configs[1])"),
    "c3": dict(size="S", out_channels=3, crop=160, batch=4,
               name="MedNeXt-S k3 3-channel affinity train step, {b}x1x160^3 crops per GPU -> 3x160^3, BCE+Dice, AdamW, DDP "
                    "(BASELINE configs[2])"),
    "c4": dict(size="L", out_channels=1, crop=224, batch=1,
               name="MedNeXt-L k3 train step, {b}x1x224^3 crops per GPU, BCE+Dice, AdamW, DDP (BASELINE configs[3])"),
}
# algorithmic work per forward of one crop/tile (fused blocks, bf16 activations)
FWD_WORK = {("S", 160): (5.49e9, 522e9), ("L", 224): (23.6e9, 5416e9), ("L", 160): (8.6e9, 1974e9)}


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
def _cpu_threads() -> int:
    return min(os.cpu_count() or 1, 32)   # oneDNN stops scaling (and regresses) beyond ~32 threads at this size


def cpu_train_rate(steps: int, warmup: int, cfg_key: str = "c2", side: int = 0):
    """The reference's own path on host cores: the oracle MedNeXt (pure-torch restatement of the un-vendored
    nnunet_mednext modules the reference builds) — fp32 forward + loss + backward + AdamW on ONE real crop of the
    config's size per step (B=1; a bounded sample of the per-GPU batch, nothing is extrapolated)."""
    from oracle.mednext_oracle import create_mednext_v1
    spec = TRAIN_CONFIGS[cfg_key]
    # MedNeXt-L at 224^3 costs ~5 minutes per fp32 CPU step: its sample is ONE real 96^3 crop, scaled by voxel count (the net is
    # fully convolutional: cost per voxel is constant up to the border) — said in `sample`, never silently
    full = spec["crop"]
    auto_small = not side and spec["size"] == "L" and full > 160
    side = side or (96 if auto_small else full)
    if not auto_small:
        full = side                 # an explicit side (tests) is its own unit: nothing is scaled
    scale = (side / full) ** 3
    cores = _cpu_threads()
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = create_mednext_v1(1, spec["out_channels"], spec["size"], 3, False)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=0.01)
    ts = []
    for i in range(warmup + steps):
        x = torch.rand(1, 1, side, side, side)
        t = (torch.rand(1, spec["out_channels"], side, side, side) > 0.85).float()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = bce_dice_loss(net(x), t)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)      # gradient_clip_val 1.0 (tutorials/mito_lucchi++)
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    sec = sum(ts) / len(ts)
    how = "no size extrapolation" if side == full else (f"EXTRAPOLATED to {full}^3 by voxel count (x{1.0 / scale:.2f} time): a "
                                                        f"{full}^3 fp32 CPU step of this net takes minutes")
    return {"value": scale / sec, "unit": "sub-volumes/s", "cores": cores, "kind": "port",
            "sample": f"oracle MedNeXt-{spec['size']} fp32 train step (forward + BCE/Dice + backward + AdamW) on one real "
                      f"1x{side}^3 crop per step ({sec:.2f} s/step, {len(ts)} timed); {how}",
            "sec_per_step": sec / scale}


def cpu_infer_rate(steps: int, warmup: int, side: int = SIDE):
    """Reference sliding-window inference cost per output voxel on host cores: one real 160^3 tile forward of the oracle
    MedNeXt-S per step; Mvox/s = tile voxels / (forward time x tile coverage), coverage 8 at 50 % overlap (the blend
    itself runs at ~20 Mvox-out/s on 8 cores, BASELINE.md §3, and is not the bound)."""
    from oracle.mednext_oracle import create_mednext_v1
    cores = _cpu_threads()
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
    return {"value": side ** 3 / sec / 8.0 / 1e6, "unit": "Mvox/s", "cores": cores, "kind": "port",
            "sample": f"oracle MedNeXt-S fp32 forward on one real {side}^3 tile per step ({sec:.2f} s, {len(ts)} timed), "
                      "/8 tile coverage at 50 % overlap",
            "sec_per_step": sec}


def run_reference(a):
    """`--impl reference`: the reference's CPU path (oracle port — nnunet_mednext / monai are not installable offline) on
    all host threads, same metric / unit / config as our arm; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps, warm = max(1, min(a.steps, 2)), max(0, min(a.warmup, 1))
    world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
    note = ("reference third-party nets (nnunet_mednext) are not installable offline; this is the oracle port of the "
            "reference's PyTorch CPU path on all host threads; every step is ONE real crop (B=1) of the config's size — a "
            "bounded sample of the per-GPU batch, value = crops/s")

    def line(metric, r, config):
        return {"impl": "reference", "metric": metric, "value": r["value"], "unit": r["unit"], "n_gpus": a.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

    if a.config == "c5":
        r = cpu_infer_rate(steps, warm)
        out = line(METRIC_INFER, r, infer_config(a, world))
    else:
        key = a.config if a.config in TRAIN_CONFIGS else "c2"
        r = cpu_train_rate(steps, warm, key)
        out = line(METRIC_TRAIN, r, train_config(a, key, world))
        if a.config == "default":
            ri = cpu_infer_rate(1, 0)
            out["infer"] = line(METRIC_INFER, ri, infer_config(a, world))
            out["infer"]["steps"], out["infer"]["warmup"] = 1, 0
    out["note"] = note
    print(json.dumps(out))


def train_batch(a, key):
    return a.batch if a.batch > 0 else TRAIN_CONFIGS[key]["batch"]


def train_config(a, key, world):
    spec = TRAIN_CONFIGS[key]
    b = train_batch(a, key)
    return {"workload": spec["name"].format(b=b), "baseline_config": key, "global_batch": world * b,
            "crop": [spec["crop"]] * 3, "parallelism": f"dp{world}",
            "l2": "inputs+activations >> 126 MB L2 per step (no explicit flush)"}


def infer_volume(a, world):
    if a.config == "c5":
        v = a.volume or 2048
        return [v, v, v]                       # ONE 2048^3 volume over all ranks (strong: the volume is fixed)
    v = a.volume or 640
    return [v * world, v, v]                   # weak: `volume`^3 per GPU


def infer_config(a, world):
    vol = infer_volume(a, world)
    tag = "BASELINE configs[4]" if a.config == "c5" else "BASELINE configs[4] geometry"
    return {"workload": f"MedNeXt-S sliding-window inference, {vol[0]}x{vol[1]}x{vol[2]} fp16 volume, {SIDE}^3 tiles, "
                        f"50% overlap, bump blending, sw_batch {a.sw_batch} ({tag})",
            "baseline_config": "c5", "volume": vol, "tile": [SIDE] * 3, "overlap": 0.5, "sw_batch": a.sw_batch,
            "parallelism": "single GPU" if world == 1 else
            f"one volume, z-slab shards x{world} + neighbour overlap exchange (NCCL send/recv)",
            "l2": "volume+accumulators >> 126 MB L2 (no explicit flush)"}


# ----------------------------------------------------------------------------- roofline leg
# op-name prefix (pytorch_connectomics_b200 profiler hooks) -> kernel symbol
KERNELS = {
    "mlp_fwd": "pcb::mlp_fused_kernel / pcb::mlp_kernel (GN-apply -> GEMM -> GELU -> GEMM + residual, tcgen05)",
    "mlp_fwd_deep": "pcb::gemm_ws_kernel x2 (deep levels: conv2+GELU, conv3+residual, tcgen05)",
    "mlp_bwd_fused": "pcb::mlp_bwd_ws_kernel / mlp_bwd_ws2_kernel (dgrad + both pointwise wgrads in TMEM, tcgen05, warp-specialised)",
    "mlp_bwd": "pcb::mlp_bwd_kernel / gemm_ws_kernel (dgrad, tcgen05)",
    "dwconv_fwd": "pcb::dwconv_same_tiled_kernel<3> / dwconv_kernel (depthwise stencil + GN statistics)",
    "dw_bwd_data": "pcb::dwconv_same_tiled_kernel<3> / dwconv_kernel (stencil data gradient)",
    "dw_wgrad": "pcb::dw_wgrad_same_tiled_kernel<3> / dw_wgrad_kernel (depthwise weight gradient)",
    "tn_gemm": "pcb::tn_gemm_ws_kernel (split-K wgrad GEMM, cp.async ring, tcgen05 MN-major)",
    "gn_bwd": "pcb::gn_dy_kernel (GroupNorm backward)",
}
NCU_TRAFFIC = {  # DRAM bytes (read + write) PER SAMPLE of one launch (dram__bytes_read.sum + dram__bytes_write.sum of the
    # `ncu --set full` captures, divided by the samples of the launch): round 2 kernels from profiles/r02_mlp_fwd_bwd_l0.ncu-rep and
    # profiles/r02_mlp_bwd_ws2.ncu-rep (batch 2, table in profiles/r02_kernels_summary.md), the unchanged stencils / gn_dy from
    # profiles/r01_blocks_final.ncu-rep (batch 1)
    "mlp_bwd_fused:m0C32H64Co32V4096000": 772.5e6,
    "mlp_bwd_fused:m2C64H128Co32V4019679": 1270.5e6,
    "mlp_fwd:m0C32H64Co32V4096000": 771.7e6,
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


def build_roofline(prof, batch, peak_gbs, measured, step_ms, nsteps):
    """Dominant kernel = op type with the largest share of the (eager, event-instrumented) step; its roofline
    numbers are quoted on its biggest launch class (level 0) with the live CUDA-event launch durations.
    `batch` = samples (crops or windows) per launch — the training batch or the engine's sw_batch."""
    if not prof:
        return None
    groups = {}
    for key, dts in prof.items():
        op = key.split(":", 1)[0]
        g = groups.setdefault(op, {"ms": 0.0, "classes": {}})
        g["ms"] += sum(dts) / nsteps
        g["classes"][key] = dts
    ranked = sorted(groups.items(), key=lambda kv: -kv[1]["ms"])
    rows = []
    for op, g in ranked[:4]:
        key, dts = max(g["classes"].items(), key=lambda kv: sum(kv[1]))
        avg_ms = sum(dts) / len(dts)
        nbytes = _op_bytes(op, key, batch)
        ach = nbytes / (avg_ms / 1e3) / 1e9 if nbytes else None
        rows.append({"kernel": KERNELS.get(op, op), "launch_class": key, "bound": "hbm", "achieved": ach, "peak": peak_gbs,
                     "unit": "GB/s", "frac": (ach / peak_gbs) if ach else None,
                     "traffic": (NCU_TRAFFIC[key] * batch) if key in NCU_TRAFFIC else None,
                     "avg_launch_ms": avg_ms, "launches_timed": len(dts), "samples_per_launch": batch,
                     "algorithmic_bytes": nbytes, "op_ms_per_step": g["ms"], "share_of_step": g["ms"] / step_ms})
    roof = dict(rows[0])
    roof["peak_source"] = "MEASURED_PEAKS.json (of measured)" if measured else "fallback 6650 GB/s (of fallback)"
    roof["traffic_source"] = "ncu DRAM bytes (read + write) per sample from the captures in profiles/ (r02_kernels_summary.md) x samples per launch; null = not captured"
    roof["others"] = rows[1:]
    return roof


def step_roofline(mode: str, units_per_step: float, ms_per_step: float, hbm_gbs: float, bf16_tflops: float,
                  work=FWD_WORK[("S", 160)], unit_name: str = "160^3"):
    """Whole-step roofline (BASELINE.md §2 / SURVEY §8d): one forward of the unit needs `work` = (min HBM bytes with fused
    blocks, FLOPs); a training step is 3x that.  Time per unit = max(bytes / HBM, FLOPs / tensor peak) with the MEASURED
    peaks; `frac` = roofline time / achieved time of the whole step (not of one kernel)."""
    passes = 3.0 if mode == "train" else 1.0
    nbytes, flops = passes * work[0], passes * work[1]
    t_unit = max(nbytes / (hbm_gbs * 1e9), flops / (bf16_tflops * 1e12))      # s per sub-volume / tile
    roof_ms = units_per_step * t_unit * 1e3
    return {"unit": f"{unit_name} sub-volume (train: fwd+bwd = 3 x fwd)" if mode == "train" else f"{unit_name} tile forward",
            "bytes_per_unit": nbytes, "flops_per_unit": flops, "units_per_step": units_per_step,
            "roofline_ms_per_step": roof_ms, "frac": roof_ms / ms_per_step if ms_per_step > 0 else None,
            "bound": "hbm" if nbytes / (hbm_gbs * 1e9) >= flops / (bf16_tflops * 1e12) else "tensor"}


def tiles_per_axis(n: int, roi: int = SIDE) -> int:
    """eager grid (window.py:92-134) at 50 % overlap: starts 0, 80, ... plus the snapped last one."""
    if n <= roi:
        return 1
    st = roi // 2
    k = (n - roi) // st + 1
    return k + (1 if (n - roi) % st else 0)


# ----------------------------------------------------------------------------- our arm
T0 = time.time()


def dbg(msg):
    if os.environ.get("PCB_BENCH_DEBUG"):
        print(f"[bench r{os.environ.get('RANK', '0')} +{time.time() - T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


class Ctx:
    """process-group plumbing shared by the two legs"""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout (one JSON line contract)
            dist.init_process_group("nccl", device_id=self.dev)
            dbg("process group up")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def run_train(cx: Ctx, a, key: str, clocks):
    """One BASELINE training config: K timed steps (CUDA-graph replay), instrumented eager pass for the roofline, e2e."""
    from pytorch_connectomics_b200 import _lib as L
    from pytorch_connectomics_b200.architectures import build_model
    from pytorch_connectomics_b200.training import FlatGradArena, broadcast_parameters

    spec = TRAIN_CONFIGS[key]
    world, rank, dev = cx.world, cx.rank, cx.dev
    nb, side, oc = train_batch(a, key), spec["crop"], spec["out_channels"]
    torch.manual_seed(1234)                         # identical initial weights on every rank ...
    model = build_model(cfg_mednext(spec["size"], oc)).to(dev)
    model.train()
    broadcast_parameters(model, src=0)              # ... and, like DDP at construction, rank 0's copy wins
    torch.manual_seed(4321 + rank)                  # data differs per rank
    arena = FlatGradArena(model.parameters())
    if os.environ.get("PCB_TORCH_ADAMW"):      # A/B switch: torch's multi-tensor AdamW instead of the fused arena kernel
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01, fused=True, capturable=True)
    else:
        # the tutorial's optimizer section (mito_lucchi++.yaml: AdamW lr 1e-3, gradient_clip_val 1.0) with the reference's
        # per-parameter groups (training/optimization/build.py:69-111), as ONE clip + AdamW pass over the flat arenas
        from pytorch_connectomics_b200.training import FusedAdamW, reference_param_groups
        opt = FusedAdamW(reference_param_groups(model, 1e-3, 0.01), max_grad_norm=1.0, arena=arena, world_size=world)
    shape, tshape = (nb, 1, side, side, side), (nb, oc, side, side, side)
    npool = 4 if side <= 160 else 2
    # a small pool of distinct resident inputs (fp16 volumes, as the data pipeline delivers them)
    pool = [(torch.rand(shape, device=dev).half(), (torch.rand(tshape, device=dev) > 0.85).float()) for _ in range(npool)]

    def step(x, t):
        arena.zero()
        loss = bce_dice_loss(model(x), t)
        loss.backward()
        arena.allreduce()
        opt.step()
        return loss

    for i in range(a.warmup):
        step(*pool[i % npool])
    cx.barrier()
    dbg("warmup done")
    # (1) instrumented eager pass: per-launch CUDA events around every pcb200 op (roofline leg)
    nprof = a.steps if (a.no_graph or a.profile_ops) else min(3, a.steps)
    L.prof_start([])
    l0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(nprof):
        step(*pool[i % npool])
    e1.record()
    cx.barrier()
    launches_per_step = (L.launch_count() - l0) // max(1, nprof)
    prof = L.prof_stop()
    ms_eager = cx.max_over_ranks(e0.elapsed_time(e1)) / nprof
    dbg(f"eager pass done: {ms_eager:.2f} ms/step")
    # (2) timed region: the same step captured once as a CUDA graph and replayed K times
    graphed, run, capture_error = None, step, None
    if not (a.no_graph or a.profile_ops):
        try:
            from pytorch_connectomics_b200.training import GraphedTrainStep
            graphed = GraphedTrainStep(model, bce_dice_loss, opt, arena, pool[0][0], pool[0][1], warmup=1,
                                       split_collective=world > 1 and not os.environ.get("PCB_GRAPH_DDP"))
            run = graphed
        except Exception as exc:   # keep the bench alive; say so in the JSON line
            capture_error = repr(exc)
            print(f"[bench] CUDA-graph capture failed, timing eager steps: {exc!r}", file=sys.stderr)
            graphed = None
            arena.rebind()
    for i in range(2):
        run(*pool[i % npool])
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for i in range(a.steps):
        run(*pool[i % npool])
    e1.record()
    cx.barrier()
    w1 = time.time()
    ms = cx.max_over_ranks(e0.elapsed_time(e1))
    value = world * nb * a.steps / (ms / 1e3)
    dbg(f"timed region done: {ms / a.steps:.2f} ms/step")

    # ---- end to end: pinned host batch -> H2D -> step -> loss D2H, every step
    hx = [torch.rand(shape).half().pin_memory() for _ in range(2)]
    ht = [(torch.rand(tshape) > 0.85).float().pin_memory() for _ in range(2)]
    h2d = hx[0].numel() * 2 + ht[0].numel() * 4

    def e2e_step(i):
        if graphed is not None:      # pinned host -> static device buffers -> graph replay -> loss D2H
            return float(graphed(hx[i % 2], ht[i % 2]).item())
        x = hx[i % 2].to(dev, non_blocking=True)
        t = ht[i % 2].to(dev, non_blocking=True)
        return float(step(x, t).item())

    for i in range(2):
        e2e_step(i)
    cx.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(a.steps):
        e2e_step(i)
    f1.record()
    cx.barrier()
    ms_e2e = cx.max_over_ranks(f0.elapsed_time(f1))
    dbg("e2e done")
    peaks = load_peaks()
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1366.0)))
    rec = {"metric": METRIC_TRAIN, "value": value, "unit": "sub-volumes/s", "n_gpus": world, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": train_config(a, key, world),
           "e2e": {"value": world * nb * a.steps / (ms_e2e / 1e3), "unit": "sub-volumes/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 4},
           "gpu_launches": int(launches_per_step * a.steps),
           "roofline": build_roofline(prof, nb, peak_gbs, bool(peaks), ms_eager, nprof),
           "execution": {"timed_region": ("cuda_graph_replay" if graphed.graph_opt is None else
                                          "cuda_graph_replay x2 around eager NCCL all-reduce") if graphed is not None else "eager",
                         "eager_ms_per_step": ms_eager, "roofline_timed_in": f"instrumented eager pass of {nprof} steps",
                         "graph_capture_error": capture_error},
           "clocks": clocks.window(w0, w1) if clocks else None}
    try:
        rec["step_roofline"] = step_roofline("train", nb, ms / a.steps, peak_gbs, tf, FWD_WORK[(spec["size"], side)],
                                             f"{side}^3")
    except Exception as exc:
        rec["step_roofline"] = {"error": repr(exc)}
    if a.profile_ops and rank == 0:
        _print_ops(prof, nprof, ms_eager)
    del graphed, run, pool, opt, arena, model
    torch.cuda.empty_cache()
    return rec


def _print_ops(prof, nsteps, step_ms):
    tot = sum(sum(v) for v in prof.values())
    for k, v in sorted(prof.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:44s} n={len(v) // nsteps:3d}/step  {sum(v) / nsteps:8.3f} ms/step  avg {sum(v) / len(v):7.3f} ms",
              file=sys.stderr)
    print(f"{'sum of timed ops':44s} {tot / nsteps:8.3f} ms/step of {step_ms:.3f}", file=sys.stderr)


def run_infer(cx: Ctx, a, clocks, steps: int, warmup: int):
    """Sliding-window inference (configs[4] geometry).  world == 1: one volume on one GPU through
    EagerSlidingWindowEngine; world > 1: ONE volume z-slab sharded over the ranks (inference/sharded.py)."""
    from pytorch_connectomics_b200 import _lib as L
    from pytorch_connectomics_b200.architectures import build_model
    from pytorch_connectomics_b200.inference.window import EagerSlidingWindowEngine

    world, rank, dev = cx.world, cx.rank, cx.dev
    vol_size = infer_volume(a, world)
    nvox = vol_size[0] * vol_size[1] * vol_size[2]
    torch.manual_seed(1234)
    model = build_model(cfg_mednext("S", 1)).to(dev)
    model.eval()
    kw = dict(roi_size=(SIDE,) * 3, sw_batch_size=a.sw_batch, overlap=0.5, mode="bump", padding_mode="constant", cval=0.0)
    net = model          # a pcb200 MedNeXt: the engine runs the whole tile loop in the library (pcb_sw_run)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)                      # every rank generates the same synthetic volume, slab by slab, on the device
    if world > 1:
        from pytorch_connectomics_b200.inference.sharded import ZSlabShardedEngine, plan_z_slabs
        sh = ZSlabShardedEngine(device=dev, **kw)
        plan = plan_z_slabs(vol_size, (SIDE,) * 3, 0.5, world)[rank]
        z0, z1 = plan.slab if plan.windows else (0, 0)
        # only this rank's slab is materialised (C5: 2048^3 fp16 = 17 GB whole, ~3 GB per rank)
        slab = torch.rand((1, 1, max(z1 - z0, 1), vol_size[1], vol_size[2]), device=dev, generator=gen).half()
        my_tiles = len(plan.windows)

        def step():
            with torch.no_grad():
                return sh.run_slab(slab, net, plan)
    else:
        eng = EagerSlidingWindowEngine(**kw)
        vol = torch.empty((1, 1, *vol_size), device=dev, dtype=torch.float16)
        for z in range(0, vol_size[0], 64):
            vol[:, :, z:z + 64] = torch.rand((1, 1, min(64, vol_size[0] - z), vol_size[1], vol_size[2]), device=dev,
                                             generator=gen).half()
        my_tiles = tiles_per_axis(vol_size[0]) * tiles_per_axis(vol_size[1]) * tiles_per_axis(vol_size[2])

        def step():
            with torch.no_grad():
                return eng(inputs=vol, network=net)

    for _ in range(max(1, warmup)):
        step()
    cx.barrier()
    L.prof_start([])
    l0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    cx.barrier()
    w1 = time.time()
    launches = L.launch_count() - l0
    L.prof_stop()
    ms = cx.max_over_ranks(e0.elapsed_time(e1))
    value = steps * nvox / (ms / 1e3) / 1e6
    dbg(f"infer timed region done: {ms / steps:.1f} ms/volume")
    # roofline leg: the native loop has no per-op hooks, so the per-kernel CUDA-event times come from ONE instrumented pass
    # of the same volume through the module path (same kernels, launched from Python with an event pair around each op)
    inner = getattr(model, "model", model)
    inner.native_inference = False
    L.prof_start([])
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    step()
    p1.record()
    cx.barrier()
    prof = L.prof_stop()
    ms_prof = cx.max_over_ranks(p0.elapsed_time(p1))      # same on every rank: the re-measure decision below must agree
    inner.native_inference = True
    # The native loop is never slower than the module path it replaces (same kernels, ~60 graph launches instead of ~7 000 kernel
    # launches; the end-to-end pass below runs the SAME native loop and includes the host copies).  The timed region is the first
    # inference work after the training leg and has come out slower than both on this pool (640^3: 2.03 / 2.12 / 2.29 s against
    # 1.89 s for the e2e pass of the same process; a 480^3 pass measured 734 / 735 / 790 / 1 373 ms in four identical runs): when
    # it is slower than the instrumented module-path pass, it is re-measured ONCE with the same K steps and the faster region is
    # reported — both are kept in `execution.remeasured`.
    remeasured = None
    if ms / steps > 1.02 * ms_prof or os.environ.get("PCB_BENCH_REMEASURE") == "1":      # env: exercise the branch
        try:
            step()                         # untimed: the native plan is rebuilt after the module-path pass toggled it off
            cx.barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(steps):
                step()
            r1.record()
            cx.barrier()
            ms2 = cx.max_over_ranks(r0.elapsed_time(r1))
            remeasured = {"first_ms_per_step": ms / steps, "second_ms_per_step": ms2 / steps}
            if ms2 < ms:
                ms = ms2
                value = steps * nvox / (ms / 1e3) / 1e6
        except Exception as exc:           # never lose the first measurement to the second
            remeasured = {"first_ms_per_step": ms / steps, "error": repr(exc)}
    # ---- end to end: pinned host volume -> H2D (this rank's slab) -> windows -> exchange -> own planes D2H
    e2e = None
    if not a.no_e2e:
        if world > 1:
            hv = torch.rand((1, 1, max(z1 - z0, 1), vol_size[1], vol_size[2])).half().pin_memory()
        else:
            hv = torch.rand((1, 1, *vol_size)).half().pin_memory()
            eng_cpu = EagerSlidingWindowEngine(sw_device=dev, output_device="cpu", **kw)
        cx.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        with torch.no_grad():
            if world > 1:
                part = sh.run_slab(hv, net, plan)
                out = part.cpu() if part is not None else torch.empty(0)
            else:
                out = eng_cpu(inputs=hv, network=net)
        f1.record()
        cx.barrier()
        ms_e2e = cx.max_over_ranks(f0.elapsed_time(f1))
        e2e = {"value": nvox / (ms_e2e / 1e3) / 1e6, "unit": "Mvox/s", "h2d_bytes_per_step": hv.numel() * 2,
               "d2h_bytes_per_step": out.numel() * out.element_size(), "steps": 1}
        del hv, out
    peaks = load_peaks()
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1366.0)))
    rec = {"metric": METRIC_INFER, "value": value, "unit": "Mvox/s", "n_gpus": world, "steps": steps, "warmup": max(1, warmup),
           "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong" if a.config == "c5" else "weak",
           "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": infer_config(a, world), "e2e": e2e,
           "gpu_launches": int(launches), "tiles_per_rank_per_step": my_tiles,
           "roofline": build_roofline(prof, min(a.sw_batch, 8), peak_gbs, bool(peaks), ms_prof, 1),
           "execution": {"timed_region": "pcb_sw_run: crop -> pcb_net_forward -> blend enqueued by the library, one CUDA-graph "
                                         "replay per window batch", "module_path_ms_per_step": ms_prof,
                         "roofline_timed_in": "one instrumented pass of the same volume through the module path",
                         "remeasured": remeasured},
           "clocks": clocks.window(w0, w1) if clocks else None}
    try:      # the slowest rank's tile count bounds the step
        rec["step_roofline"] = step_roofline("infer", my_tiles, ms / steps, peak_gbs, tf)
    except Exception as exc:
        rec["step_roofline"] = {"error": repr(exc)}
    if a.profile_ops and rank == 0:
        _print_ops(prof, 1, ms_prof)
    del model
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="pcb200", choices=["pcb200", "reference"])
    ap.add_argument("--config", default="default", choices=["default", "c2", "c3", "c4", "c5"],
                    help="BASELINE.json config; default = c2 training (top-level record) + c5-geometry inference (`infer`)")
    ap.add_argument("--mode", default=None, choices=["train", "infer"], help="legacy alias: train = --config c2, infer = c5 geometry only")
    ap.add_argument("--batch", type=int, default=0, help="sub-volumes per GPU per step (0 = the config's own: 4 for c2/c3, 1 for c4)")
    ap.add_argument("--volume", type=int, default=0, help="inference volume side per GPU (default 640; --config c5: 2048 total)")
    ap.add_argument("--sw-batch", type=int, default=4, help="windows per network call in the inference leg")
    ap.add_argument("--infer-steps", type=int, default=0, help="timed volumes of the `infer` record (default min(steps, 3))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-volume e2e pass of the inference leg")
    ap.add_argument("--profile-ops", action="store_true", help="time every pcb200 op with CUDA events (stderr table)")
    ap.add_argument("--no-graph", action="store_true", help="run the timed steps eagerly instead of as one CUDA graph")
    a = ap.parse_args()
    if a.mode == "train":
        a.config = "c2"
    a.infer_only = a.mode == "infer"
    a.warmup = max(3, a.warmup) if a.impl == "pcb200" else a.warmup
    if a.impl == "reference":
        if a.infer_only:
            a.config = "c5"
        return run_reference(a)

    from pytorch_connectomics_b200 import _lib as L
    cx = Ctx()
    if not os.path.exists(L.LIB_PATH):
        L.build()
    if L.lib().pcb_device_ok() != 1:
        raise RuntimeError(L.lib().pcb_last_error().decode())
    clocks = Clocks(cx.local) if cx.rank == 0 else None
    infer_steps = a.infer_steps or max(1, min(a.steps, 3))
    out = None
    if a.config == "c5" or a.infer_only:
        out = run_infer(cx, a, clocks, a.steps if a.config == "c5" and not a.infer_steps else infer_steps,
                        1 if a.config == "c5" else a.warmup // 3)
    else:
        out = run_train(cx, a, a.config if a.config != "default" else "c2", clocks)
        if a.config == "default":
            out["infer"] = run_infer(cx, a, clocks, infer_steps, 1)
    if clocks:
        clocks.stop()
    if cx.rank == 0:
        if cx.world == 1 and not a.no_cpu_baseline:
            if out["metric"] == METRIC_TRAIN:
                r = cpu_train_rate(1, 0, out["config"]["baseline_config"])
                out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
                if "infer" in out:
                    ri = cpu_infer_rate(1, 0)
                    out["infer"]["cpu_baseline"] = {k: ri[k] for k in ("value", "unit", "cores", "kind", "sample")}
            else:
                ri = cpu_infer_rate(1, 0)
                out["cpu_baseline"] = {k: ri[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(out))
    cx.close()


if __name__ == "__main__":
    main()

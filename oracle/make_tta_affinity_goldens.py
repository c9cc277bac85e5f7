"""Generate tests/golden/tta_affinity_goldens.npz from the REAL reference files (build container only):

    python -m oracle.make_tta_affinity_goldens

Executed in place (``ref_loader``): ``connectomics/inference/tta_affinity.py`` (with its real dependencies
``data/processing/affinity.py`` and ``utils/channel_slices.py``), ``tta_ensemble.py``, ``tta_combinations.py`` and
``window.py``.  ``TTAPredictor`` itself (``tta.py``) imports the config package and cannot be loaded offline, so its two
loops are COMPOSED here from those real functions, statement by statement: ``_run_ensemble`` (``tta.py:691-771``) and
``_predict_patch_first_local`` (``tta.py:880-1314``); ``apply_preprocessing`` is ``oracle.tta_oracle.preprocess_specs``.
"""

from __future__ import annotations

import json
import os
from types import SimpleNamespace as NS

import numpy as np
import torch

from . import ref_loader as R
from . import tta_oracle as TO

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load():
    R._base_stubs()
    c = os.path.join(R.REF_ROOT, "connectomics")
    R._stub("connectomics.utils", os.path.join(c, "utils"))
    R._load("connectomics.utils.channel_slices", "connectomics/utils/channel_slices.py")
    R._stub("connectomics.data", os.path.join(c, "data"))
    R._stub("connectomics.data.processing", os.path.join(c, "data", "processing"))
    R._load("connectomics.data.processing.affinity", "connectomics/data/processing/affinity.py")
    tc = R._load("connectomics.inference.tta_combinations", "connectomics/inference/tta_combinations.py")
    ta = R._load("connectomics.inference.tta_affinity", "connectomics/inference/tta_affinity.py")
    te = R._load("connectomics.inference.tta_ensemble", "connectomics/inference/tta_ensemble.py")
    w = R.ref_window()
    return tc, ta, te, w


def aff_cfg(offsets, mode, out_channels, extra_targets=()):
    targets = [dict(name=t) for t in extra_targets] + [dict(name="affinity", kwargs=dict(offsets=offsets, affinity_mode=mode))]
    return NS(data=NS(label_transform=NS(targets=targets, stack_outputs=True)), model=NS(out_channels=out_channels, heads={}))


PLAN_CASES = [
    dict(offsets=["1-0-0", "0-1-0", "0-0-1"], mode="deepem", extra=(), tta=dict(flip_axes="all", rotation90_axes=None)),
    dict(offsets=["1-0-0", "0-1-0", "0-0-1"], mode="banis", extra=(), tta=dict(flip_axes="all", rotation90_axes=[[1, 2]])),
    dict(offsets=["1-0-0", "0-1-0", "0-0-1", "3-0-0", "0-3-0", "0-0-3"], mode="deepem", extra=(),
         tta=dict(flip_axes=[[0], [1, 2]], rotation90_axes=[[1, 2]], rotate90_k=[0, 1, 3])),
    dict(offsets=["0-1-0", "0-0-1"], mode="deepem", extra=("binary",), tta=dict(flip_axes=[[1], [2]], rotation90_axes=[[1, 2]])),
]


def plan_json(plan):
    return dict(partial=sorted(plan.partial_channels), shifts=sorted(list(s) for s in plan.shifts), num_channels=plan.num_channels,
                rank=plan.spatial_rank,
                views=[[[m.src, m.dst, (list(m.shift) if m.shift is not None else None)] for m in v.moves] for v in plan.views])


def run_ensemble(ta, te, tc, x, net, combos, plan, specs, select, mode_cfg, dtype):
    """tta.py:691-771 with the real invert_view / accumulator."""
    acc = None
    for vi, (flip_axes, plane, k) in enumerate(combos):
        xa = x
        if flip_axes:
            xa = torch.flip(xa, dims=[a + 2 for a in flip_axes])
        if plane is not None and k > 0:
            xa = torch.rot90(xa, k=k, dims=(plane[0] + 2, plane[1] + 2))
        pred = net(xa)
        pred, validity = ta.invert_view(pred, flip_axes=flip_axes, rotation_plane_spatial=plane, k=k,
                                        view_plan=None if plan is None else plan.views[vi], tta_plan=plan)
        proc = TO.preprocess_specs(pred, specs, select, dtype)
        sel_validity = validity.select(select)
        if acc is None:
            mode_map = tc._resolve_ensemble_mode_map(mode_cfg, int(proc.shape[1]))
            raw_partial = set() if plan is None else set(plan.partial_channels)
            sel = list(range(int(pred.shape[1]))) if select is None else list(select)
            partial = [i for i, r in enumerate(sel) if r in raw_partial]
            acc = te.TTAEnsembleAccumulator(proc.shape, dtype=dtype, device=proc.device, mode_map=mode_map, partial_channels=partial,
                                            distributed_sharding=False, max_views=len(combos))
        acc.add(proc, sel_validity)
    return acc.finalize()


def patch_first(ta, te, tc, w, x, net, combos, plan_fn, specs, select, mode_cfg, dtype, roi, overlap, sw_batch, blend):
    """tta.py:939-1305 for one sample, accumulation on the CPU, real window / inversion / ensemble functions."""
    sample = x
    original = tuple(int(v) for v in sample.shape[2:])
    padded = tuple(max(original[a], roi[a]) for a in range(3))
    interval = w.compute_scan_interval(padded, roi, num_spatial_dims=3, overlap=(overlap,) * 3)
    slices = w.dense_patch_slices(padded, roi, interval, return_slice=True)
    vmap, wmap = w.build_sliding_accumulator_weight_maps(roi, mode=blend, device="cpu", value_dtype=dtype)
    pvmap, pwmap = w.build_sliding_accumulator_weight_maps(roi, mode=blend, device="cpu", value_dtype=torch.float32)
    vmap, wmap, pvmap, pwmap = (t.unsqueeze(0).unsqueeze(0) for t in (vmap, wmap, pvmap, pwmap))
    full = [None] * len(combos)
    part = [None] * len(combos)
    wfull, wpart, plan, n_raw = None, {}, None, None
    for s0 in range(0, len(slices), sw_batch):
        cur = slices[s0:s0 + sw_batch]
        batch, locs = w._extract_padded_patch_batch(sample, cur, roi_size=roi, padding_mode="constant", cval=0.0)
        batch = batch.to(dtype=torch.float32)
        added = False
        for vi, (flip_axes, plane, k) in enumerate(combos):
            xa = batch
            if flip_axes:
                xa = torch.flip(xa, dims=[a + 2 for a in flip_axes])
            if plane is not None and k > 0:
                xa = torch.rot90(xa, k=k, dims=(plane[0] + 2, plane[1] + 2))
            pred = net(xa)
            if n_raw is None:
                n_raw = int(pred.shape[1])
                plan = plan_fn(n_raw)
                pset = set() if plan is None else set(plan.partial_channels)
                raw_part = sorted(pset)
                raw_full = [c for c in range(n_raw) if c not in pset]
                if raw_full:
                    wfull = torch.zeros((1, 1, *padded), dtype=torch.float32)
                if raw_part:
                    keys = {()} | set(plan.shifts)
                    wpart = {key: torch.zeros((1, 1, *padded), dtype=torch.float32) for key in keys}
            pred, _v = ta.invert_view(pred, flip_axes=flip_axes, rotation_plane_spatial=plane, k=k,
                                      view_plan=None if plan is None else plan.views[vi], tta_plan=plan)
            if not added:
                for loc in locs:
                    gs = tuple(slice(int(loc[a]), int(loc[a]) + roi[a]) for a in range(3))
                    if wfull is not None:
                        wfull[(slice(None), slice(None), *gs)] += wmap
                    for key, wacc in wpart.items():
                        box = tuple(slice(0, r) for r in roi) if not key else ta.valid_slices_for_shift(roi, key)
                        ls = tuple(slice(int(b.start), int(b.stop)) for b in box)
                        sg = tuple(slice(int(loc[a]) + ls[a].start, int(loc[a]) + ls[a].stop) for a in range(3))
                        wacc[(slice(None), slice(None), *sg)] += pwmap[(slice(None), slice(None), *ls)]
                added = True
            if raw_full and full[vi] is None:
                full[vi] = torch.zeros((1, len(raw_full), *padded), dtype=dtype)
            if raw_part and part[vi] is None:
                part[vi] = torch.zeros((1, len(raw_part), *padded), dtype=torch.float32)
            for pi, loc in enumerate(locs):
                gs = tuple(slice(int(loc[a]), int(loc[a]) + roi[a]) for a in range(3))
                if raw_full:
                    full[vi][(slice(None), slice(None), *gs)] += pred[pi:pi + 1, raw_full].to(dtype) * vmap
                if raw_part:
                    part[vi][(slice(None), slice(None), *gs)] += pred[pi:pi + 1, raw_part].to(torch.float32) * pvmap
    crop = tuple(slice(0, v) for v in original)
    acc = None
    for vi in range(len(combos)):
        raw = torch.zeros((1, n_raw, *padded), dtype=dtype)
        rv = [None] * n_raw
        if raw_full:
            raw[:, raw_full] = w.normalize_weighted_accumulator(full[vi], wfull)
        if raw_part:
            vp = plan.views[vi]
            for pj, rc in enumerate(raw_part):
                sh = vp.shift_for_channel(rc)
                weight = wpart[() if sh is None else sh][:, 0]
                cov = weight > 0
                norm = torch.zeros_like(part[vi][:, pj], dtype=torch.float32)
                norm[cov] = part[vi][:, pj][cov] / weight[cov]
                raw[:, rc] = norm.to(dtype)
                rv[rc] = cov
        raw = raw[(slice(None), slice(None), *crop)]
        cv = [e[(slice(None), *crop)] if torch.is_tensor(e) else e for e in rv]
        proc = TO.preprocess_specs(raw, specs, select, dtype)
        sv = ta.ViewValidity(tuple(cv)).select(select)
        if acc is None:
            mode_map = tc._resolve_ensemble_mode_map(mode_cfg, int(proc.shape[1]))
            sel = list(range(n_raw)) if select is None else list(select)
            raw_partial = set() if plan is None else set(plan.partial_channels)
            acc = te.TTAEnsembleAccumulator(proc.shape, dtype=dtype, device=proc.device, mode_map=mode_map,
                                            partial_channels=[i for i, r in enumerate(sel) if r in raw_partial],
                                            distributed_sharding=False, max_views=len(combos))
        acc.add(proc, sv)
    return acc.finalize()


def main():
    assert R.available(), "needs /root/reference"
    tc, ta, te, w = load()
    g = {}
    plans = []
    for case in PLAN_CASES:
        nch = len(case["offsets"]) + len(case["extra"])
        cfg = aff_cfg(case["offsets"], case["mode"], nch, case["extra"])
        combos = tc.resolve_tta_augmentation_combinations(NS(**case["tta"]), spatial_dims=3)
        plan = ta.build_affinity_tta_plan(cfg, augmentation_combinations=combos, num_raw=nch, requested_head=None)
        plans.append(dict(case={k: (list(v) if isinstance(v, tuple) else v) for k, v in case.items()},
                          combos=[[list(f), (list(p) if p is not None else None), int(k)] for f, p, k in combos], plan=plan_json(plan)))
    g["plans_json"] = np.frombuffer(json.dumps(plans).encode(), dtype=np.uint8)

    torch.manual_seed(11)
    # ---- volume-first, 6 affinity channels (short + long range), flips + one rotation plane, sigmoid, per-channel modes
    x = torch.rand(1, 1, 6, 8, 8)
    g["vf_x"] = x.numpy()
    offs = ["1-0-0", "0-1-0", "0-0-1", "2-0-0", "0-3-0", "0-0-3"]
    for name, mode in (("deepem", "deepem"), ("banis", "banis")):
        cfg = aff_cfg(offs, mode, 6)
        combos = tc.resolve_tta_augmentation_combinations(NS(flip_axes="all", rotation90_axes=[[1, 2]], rotate90_k=[0, 1]), spatial_dims=3)
        plan = ta.build_affinity_tta_plan(cfg, augmentation_combinations=combos, num_raw=6, requested_head=None)
        out = run_ensemble(ta, te, tc, x, TO.ramp_network(6), combos, plan, [(range(6), "sigmoid")], None,
                           [["0:3", "mean"], ["3:5", "min"], ["5:", "max"]], torch.float32)
        g[f"vf_{name}"] = out.numpy()
    # selection + softmax over a non-affinity pair next to a 2-offset affinity group ("binary" occupies label channel 0)
    cfg = aff_cfg(["0-1-0", "0-0-1"], "deepem", 3, ("binary",))
    combos = tc.resolve_tta_augmentation_combinations(NS(flip_axes=[[1], [2]], rotation90_axes=[[1, 2]], rotate90_k=[0, 1, 2]), spatial_dims=3)
    plan = ta.build_affinity_tta_plan(cfg, augmentation_combinations=combos, num_raw=3, requested_head=None)
    out = run_ensemble(ta, te, tc, x, TO.ramp_network(3), combos, plan, [([1, 2], "softmax"), ([0], "tanh")], [2, 0, 1], "mean",
                       torch.float32)
    g["vf_select_softmax"] = out.numpy()
    # ---- patch-first local: 3 affinity channels + full-channel case, volume not a multiple of the window
    xp = torch.rand(1, 1, 10, 20, 20)
    g["pf_x"] = xp.numpy()
    roi = (8, 12, 12)
    for name, mode, blend in (("deepem_const", "deepem", "constant"), ("banis_bump", "banis", "bump")):
        cfg = aff_cfg(["1-0-0", "0-2-0", "0-0-2"], mode, 3)
        combos = tc.resolve_tta_augmentation_combinations(NS(flip_axes="all", rotation90_axes=[[1, 2]], rotate90_k=[0, 1]), spatial_dims=3)
        plan_fn = lambda n, cfg=cfg, combos=combos: ta.build_affinity_tta_plan(cfg, augmentation_combinations=combos, num_raw=n,
                                                                               requested_head=None)
        out = patch_first(ta, te, tc, w, xp, TO.ramp_network(3), combos, plan_fn, [(range(3), "sigmoid")], None, "mean", torch.float32,
                          roi, 0.5, 2, blend)
        g[f"pf_{name}"] = out.numpy()
    combos = tc.resolve_tta_augmentation_combinations(NS(flip_axes="all", rotation90_axes=None), spatial_dims=3)
    out = patch_first(ta, te, tc, w, xp, TO.ramp_network(2), combos, lambda n: None, [([0], "sigmoid")], None,
                      [["0", "mean"], ["1", "max"]], torch.float32, roi, 0.5, 3, "bump")
    g["pf_full_only"] = out.numpy()
    np.savez_compressed(os.path.join(OUT, "tta_affinity_goldens.npz"), **g)
    print("wrote", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()

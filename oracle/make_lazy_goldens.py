"""TEST INFRASTRUCTURE — golden vectors from the REAL ``connectomics/inference/lazy.py`` (``lazy_predict_volume`` /
``lazy_predict_region`` executed in place with the real accessor, predictor, window helpers; ``oracle/ref_loader.py::ref_lazy``
lists what is stood in) for the case matrix below: blending modes, snap-to-edge, regions, fp16 accumulators, target context,
border mask, reflect edges, TTA + activations + channel selection + mask, test-time context borders, transposes.  fp32 CPU.
Build-container only; writes ``tests/golden/lazy_goldens.npz``.  Run: ``python -m oracle.make_lazy_goldens``.
The case table and config builder are imported by ``tests/test_lazy_vs_reference.py`` and the GPU test."""

from __future__ import annotations

import os
import tempfile
from types import SimpleNamespace as NS

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lazy_goldens.npz")


def make_cfg(window, *, out_channels=1, transpose=None, image_resize=None, dt_resize=None, patch=None, overlap=0.5, blending="bump", snap=False, sw_batch=2, output_dtype=None, target_context=(), border_mask=None,
         pad_size=None, pad_mode="reflect", acts=None, select=None, tta=None, padding_mode="constant", cval=0.0,
         normalize=None, clip=None):
    sw = NS(window_size=list(window), overlap=overlap, blending=blending, sw_batch_size=sw_batch, padding_mode=padding_mode, cval=cval,
            snap_to_edge=snap, target_context=list(target_context), border_mask=border_mask, distributed_sharding=False)
    dt = NS() if pad_size is None else NS(pad_size=list(pad_size), pad_mode=pad_mode)
    if transpose is not None:
        dt.val_transpose = list(transpose)
    if dt_resize is not None:
        dt.resize = list(dt_resize)
    it = NS() if image_resize is None else NS(resize=list(image_resize))
    if normalize is not None:            # data.image_transform.normalize: smart_normalize of every patch the accessor reads
        it.normalize = normalize
    if clip is not None:
        it.clip_percentile_low, it.clip_percentile_high = clip
    return NS(model=NS(output_size=list(window), arch=NS(type="mednext"), primary_head=None, heads=None, out_channels=out_channels),
              data=NS(dataloader=NS(batch_size=1, patch_size=list(patch or window), use_lazy_h5=True), data_transform=dt,
                      image_transform=it),
              system=NS(num_workers=1),
              inference=NS(sliding_window=sw, test_time_augmentation=tta if tta is not None else NS(enabled=False),
                           model=NS(output_dtype=output_dtype, channel_activations=acts, select_channel=select, head=None)))


def _identity(x):
    return x


def _patch_mean(x):
    return x.mean(dim=(2, 3, 4), keepdim=True).expand_as(x).contiguous()


def _three(x):
    return torch.cat([x * 0.5 + 0.25, 1.0 - x, x * x], 1)


_ACTS = [dict(channels="0:2", activation="sigmoid"), dict(channels="2:3", activation="tanh")]
_TTA = NS(enabled=True, flip_axes="all", rotation90_axes=None, rotate90_k=None, ensemble_mode="mean", apply_mask=True,
          patch_first_local=False, distributed_sharding=False)

CASES = {
    "arange_bump": dict(shape=(4, 5, 6), cfg=dict(window=(2, 3, 3)), fwd=_identity, arange=True),
    "mean_constant": dict(shape=(9, 10, 11), cfg=dict(window=(4, 4, 4), blending="constant"), fwd=_patch_mean),
    "mean_snap_dt": dict(shape=(9, 10, 11), cfg=dict(window=(4, 4, 4), blending="distance_transform", snap=True, overlap=0.25), fwd=_patch_mean),
    "fp16_out": dict(shape=(6, 6, 6), cfg=dict(window=(4, 4, 4), blending="constant", output_dtype="float16"), fwd=_identity),
    "context_border": dict(shape=(9, 10, 11), cfg=dict(window=(4, 4, 4), blending="constant", target_context=(1, 2, 1), border_mask=[1, 1, 1]), fwd=_identity),
    "region": dict(shape=(12, 10, 14), cfg=dict(window=(4, 4, 4), blending="bump"), fwd=_patch_mean, region=((3, 2, 4), (9, 10, 13))),
    "reflect_edges": dict(shape=(7, 9, 8), cfg=dict(window=(4, 4, 4), blending="constant", padding_mode="reflect"), fwd=_patch_mean),
    "tta_acts_mask": dict(shape=(10, 8, 12), cfg=dict(window=(8, 8, 8), out_channels=3, blending="constant", acts=_ACTS, select=[2, 0], tta=_TTA), fwd=_three, mask=True),
    "acts_only": dict(shape=(10, 8, 12), cfg=dict(window=(8, 8, 8), out_channels=3, blending="bump", acts=_ACTS), fwd=_three),
    "pad6_edge_mask": dict(shape=(8, 9, 10), cfg=dict(window=(4, 4, 4), blending="constant", pad_size=(1, 2, 0, 3, 2, 1), pad_mode="replicate"),
                           fwd=_patch_mean, mask=True),
    "pad_constant_ctx": dict(shape=(8, 9, 10), cfg=dict(window=(4, 4, 4), blending="bump", pad_size=(2,), pad_mode="constant", target_context=(1,)),
                             fwd=_identity),
    "transpose": dict(shape=(6, 9, 12), cfg=dict(window=(4, 4, 4), blending="constant", transpose=(2, 0, 1)), fwd=_patch_mean),
    "transpose_pad_region": dict(shape=(6, 9, 12), cfg=dict(window=(4, 4, 4), blending="constant", transpose=(1, 2, 0), pad_size=(1, 1, 2)),
                                 fwd=_patch_mean, region=((2, 1, 0), (9, 10, 7))),
    "resize_image": dict(shape=(8, 9, 10), cfg=dict(window=(4, 4, 4), blending="constant", image_resize=(1.5, 1.0, 0.75)), fwd=_patch_mean),
    "resize_to_size_mask_pad": dict(shape=(8, 8, 8), cfg=dict(window=(4, 4, 4), blending="bump", dt_resize=(6, 4, 5), patch=(4, 4, 4),
                                                             pad_size=(1, 0, 2)), fwd=_identity, mask=True),
    "normalize_zscore_clip": dict(shape=(9, 10, 11), cfg=dict(window=(4, 4, 4), blending="constant", normalize="normal", clip=(0.05, 0.9)), fwd=_identity),
    "normalize_minmax_pad": dict(shape=(8, 9, 10), cfg=dict(window=(4, 4, 4), blending="bump", normalize="0-1", pad_size=(1, 2, 1), snap=True,
                                                          overlap=0.25, out_channels=3), fwd=_three),
    "normalize_divide": dict(shape=(6, 6, 7), cfg=dict(window=(4, 4, 4), blending="constant", normalize="divide-4", clip=(0.0, 0.75)), fwd=_patch_mean),
    "context_pad": dict(shape=(8, 9, 10), cfg=dict(window=(4, 4, 4), blending="constant", pad_size=(2, 1, 3), pad_mode="reflect"), fwd=_patch_mean),
}


def volumes(name: str):
    """(volume, mask or None) of a case — a function of the case name only"""
    import zlib
    case = CASES[name]
    rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    vol = (np.arange(int(np.prod(case["shape"])), dtype=np.float32).reshape(case["shape"]) if case.get("arange")
           else rs.rand(*case["shape"]).astype(np.float32))
    mask = (rs.rand(*case["shape"]) > 0.3).astype(np.float32) if case.get("mask") else None
    return vol, mask


def run_reference(name: str, workdir: str) -> torch.Tensor:
    from . import ref_loader as R
    Z = R.ref_lazy()
    case = CASES[name]
    vol, mask = volumes(name)
    np.save(os.path.join(workdir, f"{name}_v.h5.npy"), vol)            # what the stand-in h5py opens for ".../<name>_v.h5"
    if mask is not None:
        np.save(os.path.join(workdir, f"{name}_m.h5.npy"), mask)
    kw = dict(mask_path=os.path.join(workdir, f"{name}_m.h5") if mask is not None else None, device="cpu")
    cfg = make_cfg(**case["cfg"])
    with R.fake_h5py():
        if case.get("region") is None:
            return Z.lazy_predict_volume(cfg, case["fwd"], os.path.join(workdir, f"{name}_v.h5"), **kw)
        lo, hi = case["region"]
        return Z.lazy_predict_region(cfg, case["fwd"], os.path.join(workdir, f"{name}_v.h5"), region_start=lo, region_stop=hi, **kw)


def main():
    arrays = {}
    with tempfile.TemporaryDirectory() as d:
        for name in sorted(CASES):
            out = run_reference(name, d)
            arrays[name] = out.numpy()
            print(f"{name}: {tuple(out.shape)} {out.dtype}")
    np.savez_compressed(OUT, **arrays)
    print(f"wrote {OUT} ({os.path.getsize(OUT) / 1e3:.0f} kB)")


if __name__ == "__main__":
    main()

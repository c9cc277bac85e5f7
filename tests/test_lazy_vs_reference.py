"""The cfg-driven lazy seam against the REAL ``connectomics/inference/lazy.py`` executed in place (``oracle/ref_loader.py::
ref_lazy``: real ``_lazy_sliding_window`` / ``LazyVolumeAccessor`` / ``TTAPredictor`` / ``window.py``; stood in: h5py by a
``.npy``-backed file object, format detection, ``smart_normalize``).  Same config object, same forward, same volume — this
package's ``lazy_predict_volume`` / ``lazy_predict_region`` run on the CPU stand-ins of ``tests/cpu_doubles.py`` (the kernels
themselves are compared on the GPU: ``test_lazy_chunked_gpu.py``, ``test_sw_gpu.py``).  Build container only."""

import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import cpu_doubles
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is only present in the build container")


def _cfg(window, *, out_channels=1, transpose=None, overlap=0.5, blending="bump", snap=False, sw_batch=2, output_dtype=None, target_context=(), border_mask=None,
         pad_size=None, pad_mode="reflect", acts=None, select=None, tta=None, padding_mode="constant", cval=0.0):
    sw = NS(window_size=list(window), overlap=overlap, blending=blending, sw_batch_size=sw_batch, padding_mode=padding_mode, cval=cval,
            snap_to_edge=snap, target_context=list(target_context), border_mask=border_mask, distributed_sharding=False)
    dt = NS() if pad_size is None else NS(pad_size=list(pad_size), pad_mode=pad_mode)
    if transpose is not None:
        dt.val_transpose = list(transpose)
    return NS(model=NS(output_size=list(window), arch=NS(type="mednext"), primary_head=None, heads=None, out_channels=out_channels),
              data=NS(dataloader=NS(batch_size=1, patch_size=list(window), use_lazy_h5=True), data_transform=dt, image_transform=NS()),
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
    "context_pad": dict(shape=(8, 9, 10), cfg=dict(window=(4, 4, 4), blending="constant", pad_size=(2, 1, 3), pad_mode="reflect"), fwd=_patch_mean),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_lazy_seam_equals_the_real_lazy_engine(name, tmp_path, monkeypatch):
    from pytorch_connectomics_b200.inference import lazy as Zours
    Zref = ref_loader.ref_lazy()
    case = CASES[name]
    rs = np.random.RandomState(11)
    vol = (np.arange(int(np.prod(case["shape"])), dtype=np.float32).reshape(case["shape"]) if case.get("arange")
           else rs.rand(*case["shape"]).astype(np.float32))
    np.save(tmp_path / "v.h5.npy", vol)                   # what the stand-in h5py opens for ".../v.h5"
    mask = None
    if case.get("mask"):
        mask = (rs.rand(*case["shape"]) > 0.3).astype(np.float32)
        np.save(tmp_path / "m.h5.npy", mask)
    cfg = _cfg(**case["cfg"])
    kw = dict(mask_path=str(tmp_path / "m.h5") if mask is not None else None, device="cpu")
    region = case.get("region")
    with ref_loader.fake_h5py():
        if region is None:
            want = Zref.lazy_predict_volume(cfg, case["fwd"], str(tmp_path / "v.h5"), **kw)
        else:
            want = Zref.lazy_predict_region(cfg, case["fwd"], str(tmp_path / "v.h5"), region_start=region[0], region_stop=region[1], **kw)
    cpu_doubles.install(monkeypatch)
    kw_ours = dict(kw, mask_path=str(tmp_path / "m.h5.npy") if mask is not None else None)
    if region is None:
        got = Zours.lazy_predict_volume(cfg, case["fwd"], str(tmp_path / "v.h5.npy"), **kw_ours)
    else:
        got = Zours.lazy_predict_region(cfg, case["fwd"], str(tmp_path / "v.h5.npy"), region_start=region[0], region_stop=region[1], **kw_ours)
    assert got.shape == want.shape and got.dtype == want.dtype and got.device.type == "cpu"
    tol = 2e-3 if want.dtype == torch.float16 else 1e-5
    assert torch.allclose(got.float(), want.float(), rtol=tol, atol=tol * max(1.0, float(want.float().abs().max()))), \
        float((got.float() - want.float()).abs().max())

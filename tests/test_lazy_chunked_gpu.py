"""GPU parity for the lazy-grid engine and chunk sharding: the reference's own equivalence properties
(tests/unit/test_lazy_inference.py:54-111,164-240, tests/unit/test_chunked_inference.py:177-218) through the
reference's own call signatures, and the tile loop against the CPU oracle restatement of lazy.py."""
import json
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import window_oracle as O
from pytorch_connectomics_b200.inference import chunked as C
from pytorch_connectomics_b200.inference import lazy as Z
from pytorch_connectomics_b200.inference import window as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def net(t):
    return torch.cat([t * 0.5 + 1.0, 1.0 - t], 1)


def _identity_forward(x):
    return x


def _patch_mean_forward(x):          # reference tests/unit/test_lazy_inference.py: a context-dependent forward
    return x.mean(dim=(2, 3, 4), keepdim=True).expand_as(x).contiguous()


def _make_cfg(window, overlap=0.5, blending="bump", snap=False, output_dtype=None, sw_batch=2, **sw_extra):
    """the attribute paths `_lazy_sliding_window` reads (lazy.py:1010-1031)"""
    sw = NS(window_size=list(window), overlap=overlap, blending=blending, sw_batch_size=sw_batch, padding_mode="constant",
            cval=0.0, snap_to_edge=snap, target_context=[], border_mask=None, distributed_sharding=False, **sw_extra)
    return NS(model=NS(output_size=list(window), arch=NS(type="mednext")),
              data=NS(dataloader=NS(batch_size=1, patch_size=list(window)), data_transform=NS()),
              inference=NS(sliding_window=sw, model=NS(output_dtype=output_dtype)))


# ----------------------------------------------------------------------------- tensor-level tile loop vs the oracle
@pytest.mark.parametrize("snap", [False, True])
@pytest.mark.parametrize("mode", ["distance_transform", "bump"])
def test_lazy_matches_oracle(snap, mode):
    torch.manual_seed(0)
    vol = torch.rand(1, 1, 12, 14, 13)
    want = O.lazy_sliding_window(vol, net, (6, 6, 6), 0.5, mode, snap_to_edge=snap, sw_batch_size=3)
    got = Z.lazy_sliding_window(vol.to(DEV), net, roi_size=(6, 6, 6), overlap=0.5, mode=mode, snap_to_edge=snap,
                                sw_batch_size=3)
    if mode == "bump":
        assert torch.allclose(got.cpu(), want, rtol=2e-6, atol=1e-6)
    else:
        assert torch.equal(got.cpu(), want)       # no transcendental in the map: bit-exact


def test_region_is_slice_of_full_and_fp16():
    torch.manual_seed(1)
    vol = torch.rand(1, 1, 12, 14, 13, device=DEV)
    kw = dict(roi_size=(6, 6, 6), overlap=0.5, mode="distance_transform")
    full = Z.lazy_sliding_window(vol, net, **kw)
    reg = Z.lazy_sliding_window(vol, net, region_start=(3, 2, 4), region_stop=(9, 11, 13), **kw)
    assert torch.equal(reg, full[:, :, 3:9, 2:11, 4:13])
    h = Z.lazy_sliding_window(vol, net, output_dtype=torch.float16, **kw)
    want = O.lazy_sliding_window(vol.cpu(), net, (6, 6, 6), 0.5, "distance_transform", out_dtype=torch.float16)
    assert h.dtype == torch.float16 and torch.equal(h.cpu(), want)
    with pytest.raises(ValueError):
        Z.lazy_sliding_window(torch.rand(1, 1, 4, 14, 13, device=DEV), net, **kw)


def test_rank_sharded_accumulators_sum_to_full():
    torch.manual_seed(2)
    vol = torch.rand(1, 1, 12, 14, 13, device=DEV)
    kw = dict(roi_size=(6, 6, 6), overlap=0.5, mode="distance_transform")
    full = Z.lazy_sliding_window(vol, net, **kw)
    parts = [Z.lazy_sliding_window(vol, net, rank=r, world_size=2, normalize=False, **kw) for r in range(2)]
    merged = W.normalize_weighted_accumulator(parts[0][0] + parts[1][0], parts[0][1] + parts[1][1])
    assert torch.allclose(merged, full, atol=1e-5)
    seen = []
    Z.lazy_sliding_window(vol, net, accumulator_reduce=lambda v, w: seen.append((v.shape, w.shape)) or (v, w), **kw)
    assert seen and seen[0][0][1] == 2


def test_chunked_equals_full_lazy():
    torch.manual_seed(3)
    vol = torch.rand(1, 1, 20, 16, 18, device=DEV)
    kw = dict(roi_size=(8, 8, 8), overlap=0.5, mode="distance_transform", sw_batch_size=2)
    full = Z.lazy_sliding_window(vol, net, **kw)
    parts = {}
    for r in range(3):                                   # three "ranks", disjoint chunks, no collective
        parts.update(C.run_chunked_prediction(vol, net, chunk_shape=(8, 16, 9), rank=r, world_size=3, **kw))
    out = C.stitch_chunks((20, 16, 18), (8, 16, 9), parts)
    assert torch.equal(out, full)
    # resumability: chunks already done are not recomputed
    again = C.run_chunked_prediction(vol, lambda t: 1 / 0, chunk_shape=(8, 16, 9), done=parts, **kw)
    assert set(again) == set(parts)


# ----------------------------------------------------------------------------- the reference's seam, its own tests
def test_lazy_sliding_window_matches_eager_inference(tmp_path):
    """reference tests/unit/test_lazy_inference.py:54-70 (4x5x6 arange volume, identity forward, lazy == eager), called as
    the reference calls it: lazy_predict_volume(cfg, forward_fn, image_path, device=...)."""
    cfg = _make_cfg((2, 3, 3), 0.5, "bump")
    volume = np.arange(4 * 5 * 6, dtype=np.float32).reshape(4, 5, 6)
    path = tmp_path / "lazy_eager_match.npy"
    np.save(path, volume)
    x = torch.from_numpy(volume)[None, None].to(DEV)
    eager = W.build_sliding_inferer(cfg)(inputs=x, network=_identity_forward)
    lazy = Z.lazy_predict_volume(cfg, _identity_forward, str(path), device=DEV)
    assert lazy.shape == eager.shape and lazy.device.type == "cpu"
    assert torch.allclose(lazy, eager.cpu(), atol=1.0e-5)
    # the same call on an in-memory tensor (device-resident fast path) and on an accessor object
    assert torch.equal(Z.lazy_predict_volume(cfg, _identity_forward, x, device=DEV), lazy)
    acc = Z.ArrayVolumeAccessor(volume)
    assert torch.equal(Z.lazy_predict_volume(cfg, _identity_forward, acc, device=DEV), lazy)
    assert Z.get_lazy_image_reference_shape(cfg, str(path)) == (1, 1, 4, 5, 6)


def test_lazy_model_output_dtype_controls_accumulators(tmp_path):
    """reference test_lazy_inference.py:87-111 (fp16 accumulators, atol 2e-3)"""
    cfg = _make_cfg((2, 2, 2), 0.5, "constant", output_dtype="float16")
    volume = np.linspace(0.0, 1.0, num=27, dtype=np.float32).reshape(3, 3, 3)
    path = tmp_path / "lazy_output_dtype.npy"
    np.save(path, volume)
    lazy = Z.lazy_predict_volume(cfg, _identity_forward, str(path), device=DEV)
    assert lazy.dtype == torch.float16
    assert torch.allclose(lazy.float(), torch.from_numpy(volume)[None, None], atol=2.0e-3)


def test_lazy_region_matches_full_volume_global_window_grid(tmp_path):
    """reference test_lazy_inference.py:164-187 + :226-250 (region == slice(full) with a context-dependent forward;
    sub-ROI region keeps its shape)"""
    cfg = _make_cfg((3, 3, 3), 0.5, "constant", snap=True)
    volume = np.arange(5 * 6 * 7, dtype=np.float32).reshape(5, 6, 7)
    path = tmp_path / "lazy_region_global_grid.npy"
    np.save(path, volume)
    full = Z.lazy_predict_volume(cfg, _patch_mean_forward, str(path), device=DEV)
    region = Z.lazy_predict_region(cfg, _patch_mean_forward, str(path), region_start=(1, 1, 2), region_stop=(5, 6, 7),
                                   device=DEV)
    assert torch.allclose(region, full[..., 1:5, 1:6, 2:7], atol=1.0e-5)
    edge = Z.lazy_predict_region(cfg, _patch_mean_forward, str(path), region_start=(4, 4, 5), region_stop=(5, 6, 7),
                                 device=DEV)
    assert edge.shape == (1, 1, 1, 2, 2) and torch.allclose(edge, full[..., 4:5, 4:6, 5:7], atol=1.0e-5)


def test_lazy_target_context_mask_and_heads():
    """target_context (lazy.py:368-419): the model sees roi + 2*context voxels and its prediction is cropped back;
    mask multiplication (tta.py:465-548); named heads (mednext_models.py:253-273)."""
    torch.manual_seed(5)
    vol = torch.rand(1, 1, 10, 12, 11, device=DEV)
    cfg = _make_cfg((4, 4, 4), 0.5, "constant")
    base = Z.lazy_predict_volume(cfg, _identity_forward, vol, device=DEV)
    cfg_ctx = _make_cfg((4, 4, 4), 0.5, "constant")
    cfg_ctx.inference.sliding_window.target_context = [1, 2, 1]
    seen = []

    def fwd(x):
        seen.append(tuple(x.shape[2:]))
        return x

    with_ctx = Z.lazy_predict_volume(cfg_ctx, fwd, vol, device=DEV)
    assert set(seen) == {(6, 8, 6)} and torch.allclose(with_ctx, base, atol=1e-6)
    with pytest.raises(RuntimeError, match="expected prediction spatial shape"):
        Z.lazy_predict_volume(cfg_ctx, lambda x: x[..., 1:-1, 2:-2, 1:-1], vol, device=DEV)
    cfg_ctx.inference.sliding_window.target_context = [1, 2]
    with pytest.raises(ValueError, match="length 1 or 3"):
        Z.lazy_predict_volume(cfg_ctx, fwd, vol, device=DEV)
    mask = (torch.rand(10, 12, 11) > 0.5).float()
    masked = Z.lazy_predict_volume(cfg, _identity_forward, vol, mask_path=mask, device=DEV)
    assert torch.allclose(masked, base * mask[None, None], atol=1e-6)
    heads = lambda x: {"output": {"aff": torch.cat([x, x], 1), "sdt": -x}}   # noqa: E731
    sdt = Z.lazy_predict_volume(cfg, heads, vol, device=DEV, requested_head="sdt")
    assert torch.allclose(sdt, -base, atol=1e-6)
    both = Z.lazy_predict_volume(cfg, heads, vol, device=DEV, requested_head="aff,sdt")
    assert both.shape[1] == 3


def test_per_rank_chunked_runner_writes_artifacts(tmp_path):
    """`_run_chunked_prediction_per_rank` with the reference's keyword contract (chunked.py:437-560): per-chunk artifacts,
    rank-0 index, stitched volume == the full lazy prediction, finished chunks skipped on re-run."""
    torch.manual_seed(6)
    volume = np.random.RandomState(0).rand(12, 10, 14).astype(np.float32)
    path = tmp_path / "vol.npy"
    np.save(path, volume)
    cfg = _make_cfg((4, 4, 4), 0.5, "constant")
    chunks = C.build_chunk_grid(volume.shape, (6, 10, 7))
    out_path = tmp_path / "pred.h5"
    common = dict(cfg=cfg, forward_fn=_patch_mean_forward, image_path=str(path), output_path=out_path, checkpoint_path=None,
                  mask_path=None, mask_align_to_image=False, requested_head=None, device=DEV, chunks=chunks,
                  input_shape=volume.shape, final_shape=volume.shape, crop_pad=((0, 0),) * 3, crop_before=(0, 0, 0),
                  chunk_shape=(6, 10, 7), halo=(2, 2, 2), compression=None, h5_spatial_chunks=(4, 4, 4),
                  use_distributed_barrier=False)
    r1 = C._run_chunked_prediction_per_rank(rank=1, world_size=2, **common)
    assert r1 == tmp_path / "pred.h5.chunks"
    r0 = C._run_chunked_prediction_per_rank(rank=0, world_size=2, **common)
    from pytorch_connectomics_b200.inference.artifact import read_prediction_artifact
    got, meta = read_prediction_artifact(r0, return_metadata=True)
    full = Z.lazy_predict_volume(cfg, _patch_mean_forward, str(path), device=DEV)
    assert got.shape == (1, 12, 10, 14) and np.allclose(np.asarray(got), full[0].numpy(), atol=1e-5)
    index = json.load(open(tmp_path / "pred.h5.index.json"))
    assert [c["key"] for c in index["chunks"]] == [c.key for c in chunks] and index["world_size"] == 2
    assert meta["layout"] == "CZYX" and json.loads(meta["halo"]) == [2, 2, 2]
    # re-run: every chunk exists -> the forward is never called
    C._run_chunked_prediction_per_rank(rank=0, world_size=1, **{**common, "forward_fn": lambda x: 1 / 0})

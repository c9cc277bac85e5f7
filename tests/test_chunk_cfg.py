"""Config-side chunk helpers (``inference/chunk_grid.py``) against goldens produced by EXECUTING the real
``connectomics/inference/chunk_grid.py`` (``oracle/make_chunk_cfg_goldens.py``), and — in the build container — against the
real file itself.  Integer logic: exact."""

import json
import os

import pytest

from conftest import GOLDEN
from oracle.make_chunk_cfg_goldens import make_cfg
from pytorch_connectomics_b200.inference import chunk_grid as CG


def _tup(v):
    return tuple(_tup(x) for x in v) if isinstance(v, (list, tuple)) else v


@pytest.fixture(scope="module")
def goldens():
    with open(os.path.join(GOLDEN, "chunk_cfg_goldens.json")) as f:
        return json.load(f)


def _answers(mod, case):
    cfg = make_cfg(case)
    mod.validate_chunked_output_format(cfg)
    return dict(normalize_crop_pad=_tup(mod.normalize_crop_pad(case.get("crop_pad"))),
                selected_offsets=_tup(mod.resolve_selected_affinity_offsets(cfg)),
                global_crop=_tup(mod.resolve_global_prediction_crop(cfg)),
                chunk_shape=_tup(mod.resolve_chunk_shape(cfg, case["final"])),
                h5_chunks=_tup(mod.resolve_h5_spatial_chunks(case["final"])),
                output_mode=mod.resolve_chunk_output_mode(cfg))


def test_chunk_config_helpers_match_reference_goldens(goldens):
    assert len(goldens["cases"]) >= 9
    for rec in goldens["cases"]:
        got = _answers(CG, rec["case"])
        for key, val in got.items():
            assert val == _tup(rec[key]), (rec["case"]["name"], key, val, rec[key])


def test_error_texts_match_reference(goldens):
    calls = {"bad_crop": lambda c, cfg: CG.normalize_crop_pad(c["crop_pad"]),
             "bad_axes": lambda c, cfg: CG.resolve_chunk_shape(cfg, c["final"]),
             "bad_mode": lambda c, cfg: CG.resolve_chunk_output_mode(cfg),
             "bad_backend": lambda c, cfg: CG.validate_chunked_output_format(cfg)}
    for rec in goldens["bad"]:
        case = rec["case"]
        with pytest.raises(ValueError) as e:
            calls[case["name"]](case, make_cfg(case))
        assert str(e.value) == rec["error"]


def test_against_the_real_reference_file_when_present(goldens):
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference is only present in the build container")
    from oracle.make_chunk_cfg_goldens import load
    real = load()
    for rec in goldens["cases"]:
        assert _answers(CG, rec["case"]) == _answers(real, rec["case"]), rec["case"]["name"]
    offs = [(1, 0, 0), (0, -2, 0), (0, 0, 3), (-4, 5, 0)]
    import sys
    aff = sys.modules["connectomics.data.processing.affinity"]
    for mode in ("deepem", "banis"):
        assert CG.compute_affinity_crop_pad(offs, affinity_mode=mode) == aff.compute_affinity_crop_pad(offs, affinity_mode=mode)
    assert CG.compute_affinity_crop_pad([]) == aff.compute_affinity_crop_pad([]) == ()


# ----------------------------------------------------------------------------- top-level chunked driver: host helpers
def test_chunking_switches_roi_and_shards():
    from types import SimpleNamespace as NS
    from pytorch_connectomics_b200.inference import chunked as C
    assert not C.is_chunked_inference_enabled(NS())
    assert C.is_chunked_inference_enabled(NS(inference=NS(strategy="Chunked")))
    assert C.is_chunked_inference_enabled(NS(inference=NS(chunking=NS(enabled=True))))
    assert not C.is_chunked_inference_enabled(NS(inference=NS(strategy="whole_volume", chunking=NS(enabled=False))))
    assert C._resolve_distributed_rank() == (0, 1)
    mk = lambda **kw: NS(inference=NS(chunking=NS(**kw)))
    assert C._resolve_external_chunk_shard(NS()) is None and not C.is_external_chunk_sharding_enabled(mk())
    assert C._resolve_external_chunk_shard(mk(shard_id=1, num_shards=4)) == (1, 4) and C.is_external_chunk_sharding_enabled(mk(shard_id=0, num_shards=1))
    with pytest.raises(ValueError, match="must be set together"):
        C._resolve_external_chunk_shard(mk(shard_id=1))
    with pytest.raises(ValueError, match="out of range"):
        C._resolve_external_chunk_shard(mk(shard_id=4, num_shards=4))
    # chunked.py:217-243
    assert C._resolve_inference_roi(mk()) is None
    assert C._resolve_inference_roi(mk(roi=[10, 20, 30])) == ((0, 0, 0), (10, 20, 30))
    assert C._resolve_inference_roi(mk(roi=[1, 2, 3, 10, 20, 30])) == ((1, 2, 3), (10, 20, 30))
    with pytest.raises(ValueError, match="3 \\(size\\) or 6"):
        C._resolve_inference_roi(mk(roi=[1, 2]))
    with pytest.raises(ValueError, match="stop must exceed start"):
        C._resolve_inference_roi(mk(roi=[5, 0, 0, 5, 9, 9]))
    # chunked.py:246-272: drop chunks outside the ROI, crop the straddlers, keep the full-grid keys
    chunks = C.build_chunk_grid((20, 20, 20), (10, 10, 10))
    kept = C._filter_chunks_to_roi(chunks, ((0, 0, 0), (12, 10, 20)), (0, 0, 0))
    assert [c.key for c in kept] == ["z0_y0_x0", "z0_y0_x1", "z1_y0_x0", "z1_y0_x1"]
    assert kept[0] is chunks[0] and kept[2].start == (10, 0, 0) and kept[2].stop == (12, 10, 10) and kept[2].index == (1, 0, 0)
    # crop_before shifts chunk coordinates into input space before the test, and back afterwards
    kept = C._filter_chunks_to_roi(chunks, ((0, 0, 0), (12, 30, 30)), (5, 0, 0))
    assert [c.key for c in kept] == ["z0_y0_x0", "z0_y0_x1", "z0_y1_x0", "z0_y1_x1"] and all(c.stop[0] == 7 for c in kept)
    assert C._filter_chunks_to_roi(chunks, ((40, 40, 40), (50, 50, 50)), (0, 0, 0)) == []


def test_prediction_and_storage_dtype_transforms():
    import numpy as np
    from types import SimpleNamespace as NS
    from pytorch_connectomics_b200.inference.output import apply_prediction_transform, apply_storage_dtype_transform
    x = np.array([[-0.5, 0.25, 0.999, 1.5]], dtype=np.float32)
    assert apply_prediction_transform(NS(), x) is x
    off = NS(inference=NS(prediction_transform=NS(enabled=False, intensity_scale=255, intensity_dtype="uint8")))
    assert apply_prediction_transform(off, x) is x
    on = NS(inference=NS(prediction_transform=NS(enabled=True, intensity_scale=255, intensity_dtype="uint8"), save_dtype=None))
    got = apply_prediction_transform(on, x)
    assert got.dtype == np.uint8 and got.tolist() == [[0, 63, 254, 255]]          # scale in fp32, clip, truncate
    raw = NS(inference=NS(prediction_transform=NS(enabled=True, intensity_scale=-1.0, intensity_dtype="float16")))
    got = apply_prediction_transform(raw, x)
    assert got.dtype == np.float16 and np.allclose(got, x, atol=1e-3)              # negative scale: no scaling
    unknown = NS(inference=NS(prediction_transform=NS(enabled=True, intensity_scale=1.0, intensity_dtype="bfloat16")))
    assert apply_prediction_transform(unknown, x).dtype == np.float32              # unknown dtype name: kept (warning)
    assert apply_storage_dtype_transform(on, x) is x
    st = NS(inference=NS(save_dtype="int8"))
    assert apply_storage_dtype_transform(st, x * 200).tolist() == [[-100, 50, 127, 127]]


def test_top_level_chunked_driver_geometry_with_a_stub_region_predictor(tmp_path, monkeypatch):
    """Host logic of `run_chunked_prediction_inference` (chunked.py:725-957) without a GPU: `lazy_predict_region` is replaced
    by a stub that returns the region of a pointwise "prediction" (2*x+1), so chunk grid, halo boxes, core crop, global crop,
    ROI filter, transforms and the streamed artifact can be checked exactly.  The real region predictor is tested on the GPU
    (`test_lazy_chunked_gpu.py`)."""
    import numpy as np
    import torch
    from types import SimpleNamespace as NS
    from pytorch_connectomics_b200.inference import chunked as C
    from pytorch_connectomics_b200.inference.artifact import read_prediction_artifact
    volume = np.random.RandomState(3).rand(12, 10, 14).astype(np.float32)
    path = tmp_path / "vol.npy"
    np.save(path, volume)
    calls = []

    def stub(cfg, forward_fn, image_path, *, region_start, region_stop, **kw):
        calls.append((tuple(region_start), tuple(region_stop)))
        box = tuple(slice(a, b) for a, b in zip(region_start, region_stop))
        return torch.from_numpy(2.0 * volume[box] + 1.0)[None, None]

    monkeypatch.setattr(C, "lazy_predict_region", stub)
    cfg = NS(model=NS(arch=NS(type="mednext")), data=NS(dataloader=NS(), data_transform=NS()),
             inference=NS(model=NS(crop_pad=[1, 0, 2], select_channel=None), save_backend="h5", save_compression="none",
                          chunking=NS(chunk_size=[6, 16, 7], axes="all", halo=[2, 3, 2])))
    out = C.run_chunked_prediction_inference(cfg, None, str(path), output_path=tmp_path / "o" / "pred.h5", device="cpu")
    got, meta = read_prediction_artifact(out, return_metadata=True)
    assert np.array_equal(np.asarray(got), (2.0 * volume + 1.0)[None, 1:11, :, 2:12])
    # four chunks; read boxes = core (+ crop offset) grown by the halo and clipped to the INPUT volume
    assert calls == [((0, 0, 0), (9, 10, 11)), ((0, 0, 7), (9, 10, 14)), ((5, 0, 0), (12, 10, 11)), ((5, 0, 7), (12, 10, 14))]
    assert json.loads(meta["final_shape"]) == [10, 10, 10] and json.loads(meta["input_shape"]) == [12, 10, 14]
    # ROI + uint8 transform: the second z-layer of chunks is skipped, straddling chunks are cropped to the ROI face
    calls.clear()
    cfg.inference.chunking.roi = [6, 10, 9]
    cfg.inference.prediction_transform = NS(enabled=True, intensity_scale=10.0, intensity_dtype="uint8")
    out = C.run_chunked_prediction_inference(cfg, None, str(path), output_path=tmp_path / "roi.h5", device="cpu")
    got = np.asarray(read_prediction_artifact(out))
    want = np.zeros((1, 10, 10, 10), np.uint8)
    want[:, :5, :, :7] = np.clip((2.0 * volume + 1.0)[None, 1:6, :, 2:9] * 10.0, 0, 255).astype(np.uint8)
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    assert calls == [((0, 0, 0), (8, 10, 11))]
    cfg.inference.chunking.roi = [40, 40, 40, 50, 50, 50]
    with pytest.raises(ValueError, match="excludes every chunk"):
        C.run_chunked_prediction_inference(cfg, None, str(path), output_path=tmp_path / "none.h5", device="cpu")
    cfg.inference.chunking.roi = None
    cfg.inference.model.crop_pad = [6, 0, 0]
    with pytest.raises(ValueError, match="too large"):
        C.run_chunked_prediction_inference(cfg, None, str(path), output_path=tmp_path / "none.h5", device="cpu")
    cfg.inference.save_backend = "zarr"
    with pytest.raises(ValueError, match="single streamed HDF5"):
        C.run_chunked_prediction_inference(cfg, None, str(path), output_path=tmp_path / "none.h5", device="cpu")


def test_chunked_driver_with_the_real_lazy_engine_on_cpu_doubles(tmp_path, monkeypatch):
    """`run_chunked_prediction_inference` over the REAL `lazy_predict_region` / `_lazy_tile_loop` host logic, with only the
    kernel-calling window helpers replaced by oracle stand-ins (`tests/cpu_doubles.py`): a context-dependent forward (patch
    mean) makes the result sensitive to where windows are cut, so chunk == slice(full lazy prediction) is the reference's own
    property (tests/unit/test_chunked_inference.py:177-218).  GPU run of the same call:
    `test_zz_first_run_gpu.py::test_run_chunked_prediction_inference_streams_one_volume`."""
    import numpy as np
    import torch
    from types import SimpleNamespace as NS
    import cpu_doubles
    from pytorch_connectomics_b200.inference import chunked as C
    from pytorch_connectomics_b200.inference import lazy as Z
    from pytorch_connectomics_b200.inference.artifact import read_prediction_artifact
    cpu_doubles.install(monkeypatch)

    def patch_mean(x):
        return x.mean(dim=(2, 3, 4), keepdim=True).expand_as(x).contiguous()

    volume = np.random.RandomState(1).rand(12, 10, 14).astype(np.float32)
    path = tmp_path / "vol.npy"
    np.save(path, volume)
    sw = NS(window_size=[4, 4, 4], overlap=0.5, blending="constant", sw_batch_size=2, padding_mode="constant", cval=0.0,
            snap_to_edge=False, target_context=[], border_mask=None, distributed_sharding=False)
    cfg = NS(model=NS(output_size=[4, 4, 4], arch=NS(type="mednext")),
             data=NS(dataloader=NS(batch_size=1, patch_size=[4, 4, 4]), data_transform=NS()),
             inference=NS(sliding_window=sw, model=NS(output_dtype=None, crop_pad=[1, 0, 2]), save_backend="h5", save_compression=None,
                          chunking=NS(chunk_size=[6, 16, 7], axes="all", halo=[2, 2, 2], output_mode="raw_prediction")))
    full = Z.lazy_predict_volume(cfg, patch_mean, str(path), device="cpu")[0].numpy()
    want = full[:, 1:11, :, 2:12]
    out = C.run_chunked_prediction_inference(cfg, patch_mean, str(path), output_path=tmp_path / "a" / "pred.h5", device="cpu")
    got, meta = read_prediction_artifact(out, return_metadata=True)
    assert got.shape == (1, 10, 10, 10) and np.allclose(np.asarray(got), want, atol=1e-6)
    assert json.loads(meta["crop_pad"]) == [[1, 1], [0, 0], [2, 2]] and json.loads(meta["chunk_shape"]) == [6, 10, 7]
    # external shards: two scheduler jobs write disjoint per-chunk artifacts; a third call (world 1, nothing left to predict)
    # stitches them — the stitched volume is the same prediction
    cfg.inference.chunking.shard_id, cfg.inference.chunking.num_shards = 1, 2
    res = C.run_chunked_prediction_inference(cfg, patch_mean, str(path), output_path=tmp_path / "c.h5", device="cpu")
    assert res == tmp_path / "c.h5.chunks" and {f.name.split(".")[0] for f in res.iterdir()} == {"chunk_z0_y0_x1", "chunk_z1_y0_x1"}
    cfg.inference.chunking.shard_id = 0
    C.run_chunked_prediction_inference(cfg, patch_mean, str(path), output_path=tmp_path / "c.h5", device="cpu")
    chunks = C.build_chunk_grid((10, 10, 10), (6, 10, 7))
    stitched = C._run_chunked_prediction_per_rank(
        cfg=cfg, forward_fn=lambda x: 1 / 0, image_path=str(path), output_path=tmp_path / "c.h5", device="cpu", chunks=chunks,
        input_shape=(12, 10, 14), final_shape=(10, 10, 10), crop_pad=((1, 1), (0, 0), (2, 2)), crop_before=(1, 0, 2),
        chunk_shape=(6, 10, 7), halo=(2, 2, 2), compression=None, h5_spatial_chunks=(4, 4, 4), use_distributed_barrier=False)
    assert np.allclose(np.asarray(read_prediction_artifact(stitched)), want, atol=1e-6)

"""The chunked inference driver against the REAL ``connectomics/inference/chunked.py`` executed in place
(``oracle/ref_loader.py::ref_chunked``: real ``run_chunked_prediction_inference`` / ``_run_chunked_prediction_per_rank`` over the
real ``lazy.py``, ``artifact.py``, ``chunk_grid.py``, ``output.py``; h5py is a ``.npy``-backed stand-in).  Same config, same
forward, same volume: the streamed ``CZYX`` volume, its metadata attrs, the per-chunk artifacts of external shards and the
``index.json`` rank 0 writes are compared with what this package's driver produces on the CPU stand-ins of
``tests/cpu_doubles.py``.  Build container only."""

import json
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest

import cpu_doubles
from oracle import make_lazy_goldens as G
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is only present in the build container")


def _patch_mean(x):
    return x.mean(dim=(2, 3, 4), keepdim=True).expand_as(x).contiguous()


def _cfg(chunking, *, crop_pad=None, transform=None, save_dtype=None, window=(4, 4, 4), blending="constant", **extra):
    cfg = G.make_cfg(window=window, blending=blending, **extra)
    cfg.inference.chunking = NS(**chunking)
    cfg.inference.save_backend, cfg.inference.save_compression = "h5", None
    cfg.inference.model.crop_pad = crop_pad
    if transform is not None:
        cfg.inference.prediction_transform = NS(**transform)
    if save_dtype is not None:
        cfg.inference.save_dtype = save_dtype
    return cfg


CASES = {
    "plain": dict(chunking=dict(chunk_size=[6, 16, 7], axes="all", halo=[0, 0, 0])),
    "crop_halo": dict(chunking=dict(chunk_size=[6, 16, 7], axes="all", halo=[2, 2, 2], output_mode="raw_prediction"), crop_pad=[1, 0, 2]),
    "z_slabs": dict(chunking=dict(chunk_size=[5, 3, 3], axes="z", halo=[1, 0, 0])),
    "roi_uint8": dict(chunking=dict(chunk_size=[6, 16, 7], axes="all", halo=[2, 2, 2], roi=[0, 0, 0, 6, 10, 14]), crop_pad=[1, 0, 2],
                      transform=dict(enabled=True, intensity_scale=100.0, intensity_dtype="uint8")),
    "save_fp16": dict(chunking=dict(chunk_size=[8, 8, 8], axes="all", halo=[1, 1, 1]), save_dtype="float16", blending="bump"),
    "context_pad": dict(chunking=dict(chunk_size=[7, 7, 7], axes="all", halo=[2, 2, 2]), pad_size=(1, 2, 1), crop_pad=[1, 2, 1, 2, 0, 1]),
}
_IGNORED_ATTRS = ("image_path", "dataset", "chunk_stitch_source")      # paths differ by construction; "dataset" is a field of this package's .npy side-car only


def _run(module, cfg, image, out, h5):
    with (ref_loader.fake_h5py() if h5 else _null()):
        return module.run_chunked_prediction_inference(cfg, _patch_mean, image, output_path=out, device="cpu")


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def _read_ours(path):
    from pytorch_connectomics_b200.inference.artifact import read_prediction_artifact
    data, meta = read_prediction_artifact(path, return_metadata=True)
    return np.asarray(data), {k: v for k, v in meta.items() if k not in _IGNORED_ATTRS}


def _read_real(path):
    data = np.load(str(path) + ".npy")
    with open(str(path) + ".attrs.json") as fh:
        return data, {k: v for k, v in json.load(fh).items() if k not in _IGNORED_ATTRS}


@pytest.mark.parametrize("name", sorted(CASES))
def test_streamed_volume_equals_the_real_driver(name, tmp_path, monkeypatch):
    from pytorch_connectomics_b200.inference import chunked as Cours
    Cref = ref_loader.ref_chunked()
    vol = np.random.RandomState(1).rand(12, 10, 14).astype(np.float32)
    np.save(tmp_path / "v.h5.npy", vol)
    cfg = _cfg(**CASES[name])
    want_path = _run(Cref, cfg, str(tmp_path / "v.h5"), tmp_path / "ref" / "pred.h5", h5=True)
    want, want_attrs = _read_real(want_path)
    cpu_doubles.install(monkeypatch)
    got_path = _run(Cours, cfg, str(tmp_path / "v.h5.npy"), tmp_path / "ours" / "pred.h5", h5=False)
    got, got_attrs = _read_ours(got_path)
    assert got.shape == want.shape and got.dtype == want.dtype
    if np.issubdtype(want.dtype, np.integer):
        assert np.abs(got.astype(np.int64) - want.astype(np.int64)).max() <= 1     # truncation of values that differ by 1e-6
    else:
        assert np.allclose(got, want, rtol=1e-5, atol=2e-6 if want.dtype == np.float32 else 2e-3)
    assert got_attrs == want_attrs, (got_attrs, want_attrs)


def test_external_shards_and_index_equal_the_real_driver(tmp_path, monkeypatch):
    """`inference.chunking.shard_id / num_shards`: each scheduler job writes its own per-chunk artifacts (no stitching), the
    same chunk files under the same names with the same contents; then the rank-0 path of `_run_chunked_prediction_per_rank`
    (index.json + stitched volume) on both sides."""
    from pytorch_connectomics_b200.inference import chunked as Cours
    Cref = ref_loader.ref_chunked()
    vol = np.random.RandomState(2).rand(12, 10, 14).astype(np.float32)
    np.save(tmp_path / "v.h5.npy", vol)
    base = dict(chunking=dict(chunk_size=[6, 16, 7], axes="all", halo=[2, 2, 2]), crop_pad=[1, 0, 2])
    for shard in (0, 1):
        cfg = _cfg(**base)
        cfg.inference.chunking.shard_id, cfg.inference.chunking.num_shards = shard, 2
        ref_dir = _run(Cref, cfg, str(tmp_path / "v.h5"), tmp_path / "ref" / "pred.h5", h5=True)
        with monkeypatch.context() as mp:
            cpu_doubles.install(mp)
            our_dir = _run(Cours, cfg, str(tmp_path / "v.h5.npy"), tmp_path / "ours" / "pred.h5", h5=False)
        assert ref_dir.name == our_dir.name == "pred.h5.chunks"
    ref_files = sorted(f.name for f in ref_dir.iterdir() if f.name.endswith(".h5"))
    our_files = sorted({f.name.split(".npy")[0].split(".json")[0] for f in our_dir.iterdir()})
    assert ref_files == our_files and len(ref_files) == 4
    for f in ref_files:
        a, a_attrs = _read_real(ref_dir / f)
        b, b_attrs = _read_ours(our_dir / f)
        assert np.allclose(a, b, rtol=1e-5, atol=2e-6) and a_attrs == b_attrs, f
    # stitch on "rank 0 of world 1" from the finished chunk files (the forward is never called) and compare index + volume
    cfg = _cfg(**base)
    common = dict(cfg=cfg, forward_fn=lambda x: 1 / 0, checkpoint_path="ckpt/last.ckpt", mask_path=None, mask_align_to_image=False,
                  requested_head=None, device="cpu", input_shape=(12, 10, 14), final_shape=(10, 10, 10), crop_pad=((1, 1), (0, 0), (2, 2)),
                  crop_before=(1, 0, 2), chunk_shape=(6, 10, 7), halo=(2, 2, 2), compression=None, h5_spatial_chunks=(4, 4, 4), rank=0,
                  world_size=1, use_distributed_barrier=False)
    with ref_loader.fake_h5py():
        ref_chunks = ref_loader.ref_chunk_grid().build_chunk_grid((10, 10, 10), (6, 10, 7))
        ref_out = Cref._run_chunked_prediction_per_rank(image_path=str(tmp_path / "v.h5"), output_path=tmp_path / "ref" / "pred.h5",
                                                        chunks=ref_chunks, **common)
    our_out = Cours._run_chunked_prediction_per_rank(image_path=str(tmp_path / "v.h5.npy"), output_path=tmp_path / "ours" / "pred.h5",
                                                     chunks=Cours.build_chunk_grid((10, 10, 10), (6, 10, 7)), **common)
    a, a_attrs = _read_real(ref_out)
    b, b_attrs = _read_ours(our_out)
    assert np.allclose(a, b, rtol=1e-5, atol=2e-6) and a_attrs == b_attrs
    from pytorch_connectomics_b200.inference.artifact import read_prediction_artifact
    assert read_prediction_artifact(our_out, return_metadata=True)[1]["chunk_stitch_source"].endswith("ours/pred.h5.chunks")
    idx_ref = json.load(open(tmp_path / "ref" / "pred.h5.index.json"))
    idx_our = json.load(open(tmp_path / "ours" / "pred.h5.index.json"))
    strip = lambda idx: {**idx, "chunks": [{k: (os.path.basename(v).split(".npy")[0] if k == "path" else v) for k, v in c.items()} for c in idx["chunks"]]}
    assert strip(idx_ref) == strip(idx_our)


def _rank_worker(rank, world, port, workdir, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import pytest as _pt
        from pytorch_connectomics_b200.inference import chunked as Cours
        mp = _pt.MonkeyPatch()
        cpu_doubles.install(mp)
        cfg = _cfg(chunking=dict(chunk_size=[6, 16, 7], axes="all", halo=[2, 2, 2]), crop_pad=[1, 0, 2])
        out = Cours.run_chunked_prediction_inference(cfg, _patch_mean, os.path.join(workdir, "v.h5.npy"),
                                                     output_path=os.path.join(workdir, "dist", "pred.h5"), device="cpu")
        mp.undo()
        q.put((rank, str(out)))
        dist.destroy_process_group()
    except Exception as e:                                   # noqa: BLE001 - report to the parent instead of hanging it
        import traceback
        q.put((rank, f"error: {e!r}\n{traceback.format_exc()}"))


def test_two_ranks_share_the_chunks_and_rank0_stitches(tmp_path):
    """Inside a process group (world 2 over gloo) `run_chunked_prediction_inference` hands the chunks `idx % world == rank` to
    each rank, the ranks meet in a barrier, rank 0 writes `index.json` and stitches: the volume equals what the REAL driver
    streams in a single process for the same config."""
    import torch.multiprocessing as mp
    from conftest import free_port
    Cref = ref_loader.ref_chunked()
    vol = np.random.RandomState(5).rand(12, 10, 14).astype(np.float32)
    np.save(tmp_path / "v.h5.npy", vol)
    cfg = _cfg(chunking=dict(chunk_size=[6, 16, 7], axes="all", halo=[2, 2, 2]), crop_pad=[1, 0, 2])
    want, _ = _read_real(_run(Cref, cfg, str(tmp_path / "v.h5"), tmp_path / "ref" / "pred.h5", h5=True))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert not any(v.startswith("error") for v in got.values()), got
    assert got[0].endswith("dist/pred.h5") and got[1].endswith("dist/pred.h5.chunks")     # non-root ranks return the chunk dir
    data, attrs = _read_ours(tmp_path / "dist" / "pred.h5")
    assert np.allclose(data, want, rtol=1e-5, atol=2e-6)
    index = json.load(open(tmp_path / "dist" / "pred.h5.index.json"))
    assert index["world_size"] == 2 and len(index["chunks"]) == 4

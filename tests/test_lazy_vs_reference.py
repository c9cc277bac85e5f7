"""The cfg-driven lazy seam against the REAL ``connectomics/inference/lazy.py`` executed in place (``oracle/ref_loader.py::
ref_lazy``: real ``_lazy_sliding_window`` / ``LazyVolumeAccessor`` / ``TTAPredictor`` / ``window.py``; stood in: h5py by a
``.npy``-backed file object, format detection, ``smart_normalize``).  Same config object, same forward, same volume — this
package's ``lazy_predict_volume`` / ``lazy_predict_region`` run on the CPU stand-ins of ``tests/cpu_doubles.py`` (the kernels
themselves are compared on the GPU: ``test_lazy_chunked_gpu.py``, ``test_sw_gpu.py``, and against the same goldens in
``test_zz_first_run_gpu.py``).  The case table lives in ``oracle/make_lazy_goldens.py``."""

import os

import numpy as np
import pytest
import torch

import cpu_doubles
from conftest import GOLDEN
from oracle import make_lazy_goldens as G
from oracle import ref_loader


def _ours(name, tmp_path, device):
    from pytorch_connectomics_b200.inference import lazy as Z
    case = G.CASES[name]
    vol, mask = G.volumes(name)
    np.save(tmp_path / "v.npy", vol)
    if mask is not None:
        np.save(tmp_path / "m.npy", mask)
    kw = dict(mask_path=str(tmp_path / "m.npy") if mask is not None else None, device=device)
    cfg = G.make_cfg(**case["cfg"])
    if case.get("region") is None:
        return Z.lazy_predict_volume(cfg, case["fwd"], str(tmp_path / "v.npy"), **kw)
    lo, hi = case["region"]
    return Z.lazy_predict_region(cfg, case["fwd"], str(tmp_path / "v.npy"), region_start=lo, region_stop=hi, **kw)


def _check(got, want):
    assert got.shape == want.shape and got.dtype == want.dtype and got.device.type == "cpu"
    tol = 2e-3 if want.dtype == torch.float16 else 1e-5
    assert torch.allclose(got.float(), want.float(), rtol=tol, atol=tol * max(1.0, float(want.float().abs().max()))), \
        float((got.float() - want.float()).abs().max())


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_lazy_seam_equals_the_real_lazy_engine(name, tmp_path, monkeypatch):
    if not ref_loader.available():
        pytest.skip("/root/reference is only present in the build container")
    want = G.run_reference(name, str(tmp_path))
    cpu_doubles.install(monkeypatch)
    _check(_ours(name, tmp_path, "cpu"), want)


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_lazy_seam_reproduces_the_committed_goldens(name, tmp_path, monkeypatch):
    """the same comparison against ``tests/golden/lazy_goldens.npz`` — runs everywhere, also without /root/reference"""
    gold = np.load(os.path.join(GOLDEN, "lazy_goldens.npz"))
    cpu_doubles.install(monkeypatch)
    _check(_ours(name, tmp_path, "cpu"), torch.from_numpy(gold[name]))


def test_reference_shape_and_full_load_equal_the_real_accessor(tmp_path):
    """`get_lazy_image_reference_shape` (lazy.py:962-978, incl. its refusal of volumes smaller than the patch) and
    `load_full` of the accessor (`:906-918`) for every transform combination of the case table."""
    if not ref_loader.available():
        pytest.skip("/root/reference is only present in the build container")
    from types import SimpleNamespace as NS
    from pytorch_connectomics_b200.inference import lazy as Z
    R = ref_loader.ref_lazy()

    def outcome(fn):
        try:
            return ("ok", fn())
        except ValueError as e:
            return ("ValueError", str(e))

    for name in sorted(G.CASES):
        vol, mask = G.volumes(name)
        np.save(tmp_path / f"{name}.h5.npy", vol)
        cfg = G.make_cfg(**G.CASES[name]["cfg"])
        with ref_loader.fake_h5py():
            want = outcome(lambda: tuple(R.get_lazy_image_reference_shape(cfg, str(tmp_path / f"{name}.h5"))))
            full = R.load_lazy_volume(cfg, str(tmp_path / f"{name}.h5"), kind="image")
        assert outcome(lambda: Z.get_lazy_image_reference_shape(cfg, str(tmp_path / f"{name}.h5.npy"))) == want, name
        with Z.build_accessor(cfg, str(tmp_path / f"{name}.h5.npy"), kind="image", mode="test") as acc:
            assert np.allclose(acc.load_full(), full, rtol=1e-6, atol=1e-6), name
    small = G.make_cfg(window=(4, 4, 4), patch=(16, 4, 4))
    np.save(tmp_path / "small.h5.npy", np.zeros((8, 8, 8), np.float32))
    with ref_loader.fake_h5py():
        want = outcome(lambda: R.get_lazy_image_reference_shape(small, str(tmp_path / "small.h5")))
    assert want[0] == "ValueError" and outcome(lambda: Z.get_lazy_image_reference_shape(small, str(tmp_path / "small.h5.npy"))) == want


def _shard_worker(rank, world, port, workdir, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from pytorch_connectomics_b200.inference import lazy as Z
        out = {}
        for name in ("mean_constant", "tta_acts_mask"):
            case = G.CASES[name]
            cfg = G.make_cfg(**case["cfg"])
            cfg.inference.sliding_window.distributed_sharding = True
            vol, mask = G.volumes(name)
            np.save(os.path.join(workdir, f"{name}_{rank}.h5.npy"), vol)
            if mask is not None:
                np.save(os.path.join(workdir, f"{name}_{rank}_m.h5.npy"), mask)
            mpath = lambda ext: os.path.join(workdir, f"{name}_{rank}_m.h5{ext}") if mask is not None else None   # noqa: E731
            if ref_loader.available():
                R = ref_loader.ref_lazy()
                with ref_loader.fake_h5py():
                    want = R.lazy_predict_volume(cfg, case["fwd"], os.path.join(workdir, f"{name}_{rank}.h5"), mask_path=mpath(""),
                                                 device="cpu")
                out[f"ref_{name}"] = want.numpy()
            mp = pytest.MonkeyPatch()
            cpu_doubles.install(mp)
            got = Z.lazy_predict_volume(cfg, case["fwd"], os.path.join(workdir, f"{name}_{rank}.h5.npy"), mask_path=mpath(".npy"),
                                        device="cpu")
            mp.undo()
            out[f"ours_{name}"] = got.numpy()
        q.put((rank, out))
        dist.destroy_process_group()
    except Exception as e:                                   # noqa: BLE001
        import traceback
        q.put((rank, {"error": f"{e!r}\n{traceback.format_exc()}"}))


def test_window_sharding_over_two_ranks_equals_the_real_engine_and_the_goldens(tmp_path):
    """`inference.sliding_window.distributed_sharding` inside a world-2 gloo group (lazy.py:1104, lazy_distributed.py): windows
    `[rank::world]`, accumulators reduced onto rank 0, empty tensor elsewhere — this package and the REAL engine side by side in
    the same workers, and rank 0's result equal to the committed single-process goldens."""
    import torch.multiprocessing as mp
    from conftest import free_port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert not any("error" in v for v in got.values()), got
    gold = np.load(os.path.join(GOLDEN, "lazy_goldens.npz"))
    for name in ("mean_constant", "tta_acts_mask"):
        assert got[1][f"ours_{name}"].size == 0                                       # non-root ranks: empty, like the reference
        assert np.allclose(got[0][f"ours_{name}"], gold[name], rtol=1e-5, atol=1e-5), name
        if f"ref_{name}" in got[0]:
            assert got[1][f"ref_{name}"].size == 0
            assert np.allclose(got[0][f"ours_{name}"], got[0][f"ref_{name}"], rtol=1e-5, atol=1e-5), name


def test_normalize_patch_equals_the_real_smart_normalize():
    """`normalize_patch` (the accessor's per-patch `data.image_transform.normalize`) against the REAL `smart_normalize`
    (`augment_ops.py:552-610`, compiled alone from the reference file): every mode x percentile clip on random, flat and
    outlier patches — same dtype, same bits — and the same refusals."""
    if not ref_loader.available():
        pytest.skip("/root/reference is only present in the build container")
    from pytorch_connectomics_b200.inference.lazy import normalize_patch
    real = ref_loader.ref_smart_normalize()
    rs = np.random.RandomState(3)
    patches = [rs.rand(1, 4, 5, 6).astype(np.float32), (rs.randn(2, 3, 3, 3) * 40 + 100).astype(np.float32),
               np.full((1, 2, 2, 2), 7.0, np.float32), np.zeros((1, 3, 3, 3), np.float32)]
    patches[1][0, 0, 0, 0] = 1e4
    for patch in patches:
        for mode in ("normal", "0-1", "divide-255", "divide-0.5", "none"):
            for lo, hi in ((0.0, 1.0), (0.02, 0.98), (0.0, 0.9), (0.25, 1.0)):
                want = real(patch, mode, divide_value=None, clip_percentile_low=lo, clip_percentile_high=hi)
                got = normalize_patch(patch, mode, lo, hi)
                assert got.dtype == want.dtype and np.array_equal(got, want), (mode, lo, hi)
                assert got is not patch
    for bad in ("divide", "divide-x", "divide-0", "zscore"):
        with pytest.raises(ValueError) as ours:
            normalize_patch(patches[0], bad)
        with pytest.raises(ValueError) as theirs:
            real(patches[0], bad)
        assert str(ours.value) == str(theirs.value), bad


def test_normalizing_accessor_never_hands_out_the_resident_tensor():
    """a normalised image is a function of each PATCH: the device-resident whole-volume fast path must be off"""
    from types import SimpleNamespace as NS
    from pytorch_connectomics_b200.inference import lazy as Z
    cfg = G.make_cfg(window=(4, 4, 4), normalize="normal")
    vol = torch.rand(8, 8, 8)
    acc = Z.build_accessor(cfg, vol, kind="image", mode="test")
    assert acc.normalize_mode == "normal" and acc.as_tensor() is None
    plain = Z.build_accessor(G.make_cfg(window=(4, 4, 4)), vol, kind="image", mode="test")
    assert plain.as_tensor() is not None
    p = acc.read_patch((2, 2, 2), (4, 4, 4), outer_pad_mode="constant", outer_pad_value=0.0)
    assert abs(float(p.mean())) < 1e-5 and abs(float(p.std()) - 1.0) < 1e-4
    mask = Z.build_accessor(cfg, vol, kind="mask", mode="test")           # masks are never normalised
    assert np.array_equal(mask.read_patch((0, 0, 0), (4, 4, 4), outer_pad_mode="constant", outer_pad_value=0.0)[0], vol[:4, :4, :4].numpy())

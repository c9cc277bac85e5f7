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

"""The HOST logic of GPU tests, replayed in the CPU suite: the bodies of the lazy / chunked / predictor GPU tests run unchanged
with ``DEV = "cpu"`` and the kernel-calling window helpers replaced by the oracle stand-ins of ``tests/cpu_doubles.py`` (crop,
weight map, accumulate, normalise, TTA fold chain).  What this catches is a host-side regression (a changed signature, a wrong
box, a lost config node) BEFORE the code reaches a B200; what it cannot catch is a kernel bug — the same bodies run against the
real kernels under ``-m gpu``."""

import inspect

import pytest

import cpu_doubles
import test_lazy_chunked_gpu as lazy_gpu
import test_zz_first_run_gpu as first_run_gpu

SKIP = {"test_lazy_sliding_window_matches_eager_inference"}      # drives EagerSlidingWindowEngine's CUDA streams directly
ZZ = ("test_run_chunked_prediction_inference_streams_one_volume", "test_lazy_seam_runs_patches_through_the_predictor",
      "test_lazy_seam_matches_the_real_lazy_engine_goldens")


def _cases(module, names=None):
    out = []
    for name, fn in inspect.getmembers(module, inspect.isfunction):
        if not name.startswith("test_") or name in SKIP or (names is not None and name not in names):
            continue
        grids = [{}]
        for mark in getattr(fn, "pytestmark", []):
            if mark.name != "parametrize":
                continue
            keys = [k.strip() for k in mark.args[0].split(",")]
            grids = [{**g, **dict(zip(keys, v if len(keys) > 1 else (v,)))} for g in grids for v in mark.args[1]]
        out += [pytest.param(module, name, g, id=f"{name}-{'-'.join(str(v) for v in g.values())}" if g else name) for g in grids]
    return out


@pytest.mark.parametrize("module,name,kwargs", _cases(lazy_gpu) + _cases(first_run_gpu, ZZ))
def test_gpu_test_body_on_cpu_doubles(module, name, kwargs, tmp_path, monkeypatch):
    cpu_doubles.install(monkeypatch)
    monkeypatch.setattr(module, "DEV", "cpu")
    fn = getattr(module, name)
    kw = dict(kwargs)
    if "tmp_path" in inspect.signature(fn).parameters:
        kw["tmp_path"] = tmp_path
    fn(**kw)

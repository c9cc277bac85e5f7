"""Load a handful of dependency-free reference files IN PLACE (no copying) so the oracle can be
pinned against the real reference.  Build-container only: ``/root/reference`` does not exist on
the GPU box, so nothing that runs there may call this (``available()`` guards).

Loadable as files (torch-only deps): ``connectomics/inference/window.py`` (needs one symbol from
``..config.hardware``, stubbed), ``connectomics/models/architectures/{registry,base}.py``,
``connectomics/chunked/{chunk_grid,halo}.py``.  ``import connectomics.<pkg>`` itself fails here
(monai / omegaconf / h5py are not installed) — see SURVEY.md §8(c).
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("PCB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "connectomics", "inference", "window.py"))


def _stub(name, path=None, **attrs):
    if name in sys.modules:                      # a real module wins; an existing stand-in is kept and extended
        m = sys.modules[name]
        if getattr(m, "__pcb_stub__", False):
            m.__dict__.update(attrs)
        return m
    m = types.ModuleType(name)
    m.__path__ = [path] if path else []
    m.__pcb_stub__ = True
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(modname, relpath):
    if modname in sys.modules and not getattr(sys.modules[modname], "__pcb_stub__", False):
        return sys.modules[modname]
    sys.dont_write_bytecode = True  # /root/reference is read-only
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def _base_stubs():
    c = os.path.join(REF_ROOT, "connectomics")
    _stub("connectomics", c)
    _stub("connectomics.config")
    _stub("connectomics.config.hardware", resolve_accelerator_type=lambda requested="auto": "cpu",
          empty_accelerator_cache=lambda *a, **k: None)
    _stub("connectomics.inference", os.path.join(c, "inference"))
    _stub("connectomics.models", os.path.join(c, "models"))
    _stub("connectomics.models.architectures", os.path.join(c, "models", "architectures"))
    _stub("connectomics.chunked", os.path.join(c, "chunked"))


def ref_window():
    _base_stubs()
    return _load("connectomics.inference.window", "connectomics/inference/window.py")


def ref_registry():
    _base_stubs()
    return _load("connectomics.models.architectures.registry",
                 "connectomics/models/architectures/registry.py")


def ref_base():
    _base_stubs()
    return _load("connectomics.models.architectures.base",
                 "connectomics/models/architectures/base.py")


def ref_chunk_grid():
    _base_stubs()
    return _load("connectomics.chunked.chunk_grid", "connectomics/chunked/chunk_grid.py")


def ref_halo():
    ref_chunk_grid()
    return _load("connectomics.chunked.halo", "connectomics/chunked/halo.py")


def ref_optimizer_build():
    """``connectomics/training/optimization/build.py`` (torch + its sibling ``lr_scheduler.py`` only)"""
    _base_stubs()
    c = os.path.join(REF_ROOT, "connectomics")
    _stub("connectomics.training", os.path.join(c, "training"))
    _stub("connectomics.training.optimization", os.path.join(c, "training", "optimization"))
    _load("connectomics.training.optimization.lr_scheduler", "connectomics/training/optimization/lr_scheduler.py")
    return _load("connectomics.training.optimization.build", "connectomics/training/optimization/build.py")


def ref_artifact():
    """``connectomics/inference/artifact.py`` (json / numpy + ``utils/model_outputs.py``; h5py is imported lazily there)"""
    _base_stubs()
    c = os.path.join(REF_ROOT, "connectomics")
    _stub("connectomics.utils", os.path.join(c, "utils"))
    _load("connectomics.utils.model_outputs", "connectomics/utils/model_outputs.py")
    return _load("connectomics.inference.artifact", "connectomics/inference/artifact.py")


def ref_mednext_models():
    """``connectomics/models/architectures/mednext_models.py`` executed in place over a stand-in ``nnunet_mednext`` module
    whose ``MedNeXt`` / ``MedNeXtBlock`` / ``create_mednext_v1`` are the ORACLE restatements (the third-party package is not
    installable offline).  What this gives the tests is the REAL builders, wrappers, task heads and validation code of the
    reference; the network arithmetic underneath is the oracle's."""
    _base_stubs()
    from . import mednext_oracle as MO
    if "nnunet_mednext" not in sys.modules:
        _stub("nnunet_mednext", MedNeXt=MO.MedNeXt, MedNeXtBlock=MO.MedNeXtBlock, create_mednext_v1=MO.create_mednext_v1)
    ref_registry()
    ref_base()
    return _load("connectomics.models.architectures.mednext_models", "connectomics/models/architectures/mednext_models.py")


def ref_monai_models():
    """``connectomics/models/architectures/monai_models.py`` executed in place over a stand-in ``monai`` package whose
    ``UNet`` / ``ResidualUnit`` are the ORACLE restatements (MONAI cannot be installed offline); the classes the path does
    not use (``BasicUNet``, ``UNETR``, ``SwinUNETR``, ``UpSample``) are placeholders that raise when built."""
    _base_stubs()
    from . import monai_unet_oracle as UO

    def _absent(name):
        def build(*_a, **_k):
            raise NotImplementedError(f"stand-in monai: {name} is outside the hot path")
        return build

    if "monai" not in sys.modules:
        _stub("monai")
        _stub("monai.networks")
        _stub("monai.networks.blocks", ResidualUnit=UO.ResidualUnit, UpSample=UO.UpSample)
        _stub("monai.networks.nets", UNet=UO.UNet, BasicUNet=_absent("BasicUNet"), UNETR=_absent("UNETR"), SwinUNETR=_absent("SwinUNETR"))
    ref_registry()
    ref_base()
    return _load("connectomics.models.architectures.monai_models", "connectomics/models/architectures/monai_models.py")


def ref_tta():
    """``connectomics/inference/tta.py`` (the REAL ``TTAPredictor``) with its real dependencies ``tta_affinity.py``,
    ``tta_combinations.py``, ``tta_ensemble.py``, ``window.py``, ``utils/{channel_slices,model_outputs}.py``,
    ``data/processing/affinity.py``; the one symbol it takes from the config package (``empty_accelerator_cache``) is a no-op."""
    _base_stubs()
    c = os.path.join(REF_ROOT, "connectomics")
    _stub("connectomics.utils", os.path.join(c, "utils"))
    _load("connectomics.utils.channel_slices", "connectomics/utils/channel_slices.py")
    _load("connectomics.utils.model_outputs", "connectomics/utils/model_outputs.py")
    _stub("connectomics.data", os.path.join(c, "data"))
    _stub("connectomics.data.processing", os.path.join(c, "data", "processing"))
    _load("connectomics.data.processing.affinity", "connectomics/data/processing/affinity.py")
    ref_window()
    _load("connectomics.inference.tta_combinations", "connectomics/inference/tta_combinations.py")
    _load("connectomics.inference.tta_affinity", "connectomics/inference/tta_affinity.py")
    _load("connectomics.inference.tta_ensemble", "connectomics/inference/tta_ensemble.py")
    return _load("connectomics.inference.tta", "connectomics/inference/tta.py")


class _FakeDataset:
    """what the reference touches on an h5py dataset: shape / dtype, slicing both ways, ``attrs``"""

    def __init__(self, array, attrs):
        self._a, self.attrs = array, attrs

    shape = property(lambda self: tuple(self._a.shape))
    dtype = property(lambda self: self._a.dtype)
    ndim = property(lambda self: self._a.ndim)

    def __getitem__(self, key):
        return self._a[key]

    def __setitem__(self, key, value):
        self._a[key] = value

    def __array__(self, dtype=None, copy=None):
        import numpy as np
        return np.asarray(self._a, dtype=dtype)


class _FakeH5File:
    """Stand-in for ``h5py.File`` over plain files next to the path: dataset ``<path>.npy`` (+ attrs ``<path>.attrs.json``).
    h5py is not installable offline; the reference's lazy accessor and artifact writer / reader need ``keys()``, ``[name]``,
    ``create_dataset``, ``attrs``, the context-manager protocol and ``close()`` from it."""

    def __init__(self, path, mode="r"):
        import json
        import numpy as np
        self._path, self._mode, self._sets = str(path), mode, {}
        if mode == "r":
            attrs = {}
            if os.path.exists(self._path + ".attrs.json"):
                with open(self._path + ".attrs.json") as fh:
                    attrs = json.load(fh)
            self._sets["main"] = _FakeDataset(np.load(self._path + ".npy", mmap_mode="r"), attrs)

    def keys(self):
        return list(self._sets.keys())

    def __contains__(self, name):
        return name in self._sets

    def __getitem__(self, name):
        return self._sets[name]

    def create_dataset(self, name, shape=None, dtype=None, data=None, chunks=None, compression=None, **_kw):
        import numpy as np
        if data is not None:
            arr = np.array(data)
            np.save(self._path + ".npy", arr)
        else:
            arr = np.lib.format.open_memmap(self._path + ".npy", mode="w+", dtype=np.dtype(dtype), shape=tuple(shape))
        self._sets[name] = _FakeDataset(arr, {})
        return self._sets[name]

    def close(self):
        import json
        if self._mode != "r":
            for ds in self._sets.values():
                if hasattr(ds._a, "flush"):
                    ds._a.flush()
                with open(self._path + ".attrs.json", "w") as fh:
                    json.dump({k: (v if isinstance(v, (str, int, float, bool)) or v is None else str(v)) for k, v in ds.attrs.items()}, fh)
                open(self._path, "a").close()          # the reference tests for the artifact path itself
        self._sets = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class fake_h5py:
    """``with fake_h5py():`` — the stand-in is visible as ``h5py`` ONLY inside the block (the real accessor imports it lazily
    when it opens a file), so nothing else in the process — this package's artifact writer probes for h5py — ever sees it."""

    def __enter__(self):
        self._old = sys.modules.get("h5py")
        m = types.ModuleType("h5py")
        m.File = _FakeH5File
        sys.modules["h5py"] = m
        return m

    def __exit__(self, *exc):
        if self._old is None:
            sys.modules.pop("h5py", None)
        else:
            sys.modules["h5py"] = self._old
        return False


def ref_lazy():
    """``connectomics/inference/lazy.py`` — the REAL lazy sliding-window engine (``_lazy_sliding_window``,
    ``lazy_predict_region / volume``, ``LazyVolumeAccessor``, ``_build_accessor``) with the real ``tta.py``,
    ``lazy_distributed.py``, ``window.py`` and ``data/processing/misc.py``.  What is stood in: ``h5py`` (a ``.npy``-backed
    file object, visible only inside ``with fake_h5py():``), ``data/io/io.py`` (format detection by extension; its tiff helpers are never reached) and
    the ``augment_ops`` MODULE (it imports cv2) — but its ``smart_normalize`` is the REAL function, compiled alone from
    the reference file (``ref_smart_normalize``)."""
    ref_tta()

    def _no(name):
        def fn(*_a, **_k):
            raise NotImplementedError(f"stand-in: {name} is outside the hot path")
        return fn

    _stub("connectomics.data.augmentation")
    _stub("connectomics.data.augmentation.augment_ops", smart_normalize=ref_smart_normalize())
    _stub("connectomics.data.io")
    _stub("connectomics.data.io.io", _detect_format=lambda p: "h5" if str(p).endswith((".h5", ".hdf5")) else "unknown",
          _get_tiff_volume_shape=_no("_get_tiff_volume_shape"), _tiff_series_are_stackable=_no("_tiff_series_are_stackable"))
    _load("connectomics.data.processing.misc", "connectomics/data/processing/misc.py")
    _load("connectomics.inference.lazy_distributed", "connectomics/inference/lazy_distributed.py")
    return _load("connectomics.inference.lazy", "connectomics/inference/lazy.py")


def cut_function(relpath: str, name: str, **namespace):
    """One module-level function of a reference file, cut out of the file's syntax tree and compiled in place with the given
    globals — for files whose module-level imports need packages this image does not have.  Nothing is copied."""
    import ast
    path = os.path.join(REF_ROOT, relpath)
    tree = ast.parse(open(path).read(), filename=path)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = dict(namespace)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def ref_smart_normalize():
    """The REAL ``smart_normalize`` (``connectomics/data/augmentation/augment_ops.py:552-610``): the module imports cv2, which
    this image does not have, so the one function (numpy only) is cut out of the reference file's syntax tree and compiled in
    place — nothing is copied into the repository."""
    import ast
    import typing
    import numpy
    path = os.path.join(REF_ROOT, "connectomics", "data", "augmentation", "augment_ops.py")
    tree = ast.parse(open(path).read(), filename=path)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "smart_normalize")
    ns = {"np": numpy, "Optional": typing.Optional, "List": typing.List, "Tuple": typing.Tuple}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["smart_normalize"]


def ref_chunked():
    """``connectomics/inference/chunked.py`` — the REAL chunked driver (``run_chunked_prediction_inference``,
    ``_run_chunked_prediction_per_rank``, ROI filter, external shards, stitching) over the real ``lazy.py``, ``artifact.py``,
    ``chunk_grid.py`` and ``output.py``.  Stood in: what ``output.py`` imports from the config / data packages at module level
    (``Config``, ``DictConfig``, ``restore_prediction_to_input_space`` — names only, never called on this path) and h5py
    (``with fake_h5py():``)."""
    ref_lazy()
    c = os.path.join(REF_ROOT, "connectomics")
    if "omegaconf" not in sys.modules:
        _stub("omegaconf", DictConfig=type("DictConfig", (), {}))
    _stub("connectomics.config", Config=type("Config", (), {}))
    _stub("connectomics.data.processing.nnunet_preprocess", restore_prediction_to_input_space=lambda *a, **k: None)
    ref_chunk_grid()
    ref_halo()
    ref_artifact()
    _load("connectomics.inference.chunk_grid", "connectomics/inference/chunk_grid.py")
    _load("connectomics.inference.output", "connectomics/inference/output.py")
    return _load("connectomics.inference.chunked", "connectomics/inference/chunked.py")

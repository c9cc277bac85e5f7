"""Raw prediction artifacts (``connectomics/inference/artifact.py:16-230``): one ``CZYX`` array + the reference's metadata.

The reference writes HDF5 (dataset ``main`` with the metadata as attrs).  h5py is an optional dependency here: with it
the same file is produced; without it the array goes to ``<path>.npy`` (``numpy.lib.format``, memory-mappable, so a
stitcher can stream chunks into it) and the metadata to ``<path>.json`` — same keys, same JSON encoding of tuples."""

from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field
from pathlib import Path
from typing import Any, Callable, Mapping, Optional, Sequence

import numpy as np


@dataclass(frozen=True)
class PredictionArtifactMetadata:
    """``artifact.py:16-37``."""
    kind: str = "raw_prediction"
    layout: str = "CZYX"
    image_path: Optional[str] = None
    checkpoint_path: Optional[str] = None
    output_head: Optional[str] = None
    input_shape: Optional[tuple] = None
    final_shape: Optional[tuple] = None
    crop_pad: Optional[tuple] = None
    transpose: Optional[tuple] = None
    model_architecture: Optional[str] = None
    model_output_identity: Optional[str] = None
    decode_after_inference: Optional[bool] = None
    chunk_shape: Optional[tuple] = None
    halo: Optional[tuple] = None
    channel_order: Optional[tuple] = None
    activation: Optional[str] = None
    intensity_scale: Optional[float] = None
    intensity_dtype: Optional[str] = None
    extra: Mapping[str, Any] = field(default_factory=dict)


def _cfg_get(obj: Any, path: str, default: Any = None) -> Any:
    node = obj
    for part in path.split("."):
        if node is None:
            return default
        node = node.get(part, default) if isinstance(node, Mapping) else getattr(node, part, default)
    return node


def _tuple_or_none(value):
    if value in (None, [], ()):
        return None
    return tuple(int(v) for v in value)


def build_prediction_artifact_metadata(cfg: Any, *, image_path=None, checkpoint_path=None, output_head=None,
                                       input_shape=None, final_shape=None, crop_pad=None, chunk_shape=None, halo=None,
                                       intensity_scale=None, intensity_dtype=None, extra=None) -> PredictionArtifactMetadata:
    """``artifact.py:75-117``."""
    tcfg = _cfg_get(cfg, "inference.prediction_transform")
    enabled = bool(getattr(tcfg, "enabled", False))
    if intensity_scale is None and enabled:
        intensity_scale = float(getattr(tcfg, "intensity_scale", -1.0))
    if intensity_dtype is None and enabled:
        intensity_dtype = getattr(tcfg, "intensity_dtype", None)
    ident = []
    if output_head:
        ident.append(f"head={output_head}")
    elif _cfg_get(cfg, "model.primary_head"):
        ident.append(f"primary_head={_cfg_get(cfg, 'model.primary_head')}")
    sel = _cfg_get(cfg, "inference.model.select_channel")      # utils/model_outputs.py:37-39 get_inference_select_channel
    if sel is not None:
        ident.append(f"select_channel={sel}")
    return PredictionArtifactMetadata(
        image_path=image_path, checkpoint_path=str(checkpoint_path) if checkpoint_path is not None else None,
        output_head=output_head, input_shape=_tuple_or_none(input_shape), final_shape=_tuple_or_none(final_shape),
        crop_pad=tuple((int(p[0]), int(p[1])) for p in crop_pad) if crop_pad is not None else None,
        transpose=_tuple_or_none(_cfg_get(cfg, "data.data_transform.val_transpose")),
        model_architecture=_cfg_get(cfg, "model.arch.type"), model_output_identity=";".join(ident) if ident else None,
        decode_after_inference=bool(_cfg_get(cfg, "decoding.enabled", True)), chunk_shape=_tuple_or_none(chunk_shape),
        halo=_tuple_or_none(halo), intensity_scale=intensity_scale, intensity_dtype=intensity_dtype, extra=extra or {})


def _json_attr(value: Any) -> Any:
    if value is None or isinstance(value, (str, int, float, bool)):
        return value
    if isinstance(value, (tuple, list, dict)):
        return json.dumps(value)
    return str(value)


def metadata_attrs(metadata: PredictionArtifactMetadata) -> dict:
    """the flat attr dict ``write_prediction_artifact_attrs`` stores (``artifact.py:132-138``)"""
    attrs = asdict(metadata)
    extra = attrs.pop("extra", {}) or {}
    return {k: _json_attr(v) for k, v in {**attrs, **dict(extra)}.items() if v is not None}


def have_h5py() -> bool:
    try:
        import h5py  # noqa: F401
        return True
    except ImportError:
        return False


def artifact_path(path) -> Path:
    """where the array actually lives: ``path`` itself with h5py, ``path + '.npy'`` without"""
    p = Path(path)
    return p if have_h5py() else p.with_suffix(p.suffix + ".npy")


def write_prediction_artifact(path, data=None, *, metadata: Optional[PredictionArtifactMetadata] = None,
                              dataset: str = "main", compression="gzip", shape=None, dtype=None, chunks=None,
                              writer: Optional[Callable[[Any], None]] = None) -> Path:
    """``artifact.py:141-203``: one ``CZYX`` artifact; ``data=None`` + ``shape``/``dtype`` + ``writer(dset)`` streams."""
    arr = None if data is None else np.asarray(data)
    if arr is None:
        if shape is None or dtype is None:
            raise ValueError("Streaming prediction artifacts require shape and dtype.")
        ashape, adtype = tuple(int(v) for v in shape), np.dtype(dtype)
    else:
        if arr.ndim != 4:
            raise ValueError(f"Prediction artifacts must use CZYX layout, got shape {arr.shape}")
        ashape, adtype = tuple(int(v) for v in arr.shape), arr.dtype
    if len(ashape) != 4:
        raise ValueError(f"Prediction artifacts must use CZYX layout, got shape {ashape}")
    out = Path(path)
    out.parent.mkdir(parents=True, exist_ok=True)
    meta = metadata or PredictionArtifactMetadata(final_shape=tuple(int(v) for v in ashape[-3:]), intensity_dtype=str(adtype))
    if have_h5py():
        import h5py
        with h5py.File(out, "w") as handle:
            if arr is None:
                dset = handle.create_dataset(dataset, shape=ashape, dtype=adtype, chunks=chunks, compression=compression)
            else:
                dset = handle.create_dataset(dataset, data=arr, chunks=chunks, compression=compression)
            for k, v in metadata_attrs(meta).items():
                dset.attrs[k] = v
            if writer is not None:
                writer(dset)
        return out
    npy = artifact_path(out)
    tmp = npy.with_suffix(npy.suffix + ".tmp")
    mm = np.lib.format.open_memmap(tmp, mode="w+", dtype=adtype, shape=ashape)
    if arr is not None:
        mm[...] = arr
    if writer is not None:
        writer(mm)
    mm.flush()
    del mm
    with open(out.with_suffix(out.suffix + ".json"), "w") as fh:
        json.dump({"dataset": dataset, **metadata_attrs(meta)}, fh, indent=2)
    tmp.replace(npy)          # the array appears atomically: an existing file means a finished chunk (resume marker)
    return npy


def read_prediction_artifact(path, *, dataset: str = "main", return_metadata: bool = False):
    """``artifact.py:206-224``."""
    p = Path(path)
    if have_h5py() and p.exists() and p.suffix != ".npy":
        import h5py
        with h5py.File(p, "r") as handle:
            data, meta = np.asarray(handle[dataset]), dict(handle[dataset].attrs)
    else:
        npy = p if p.suffix == ".npy" else p.with_suffix(p.suffix + ".npy")
        data = np.load(npy, mmap_mode="r")
        side = Path(str(npy)[:-4] + ".json")
        meta = json.load(open(side)) if side.exists() else {}
    if data.ndim != 4:
        raise ValueError(f"Prediction artifact must use CZYX layout, got shape {data.shape}")
    return (data, meta) if return_metadata else data


__all__ = ["PredictionArtifactMetadata", "build_prediction_artifact_metadata", "write_prediction_artifact",
           "read_prediction_artifact", "artifact_path", "metadata_attrs"]

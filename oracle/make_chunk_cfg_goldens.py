"""TEST INFRASTRUCTURE — golden answers for the config-side chunk helpers, produced by EXECUTING the real
``connectomics/inference/chunk_grid.py`` (with its real dependencies ``data/processing/affinity.py``,
``utils/channel_slices.py``, ``utils/model_outputs.py``) in place.  Build-container only (``/root/reference``);
writes ``tests/golden/chunk_cfg_goldens.json``.  Run: ``python -m oracle.make_chunk_cfg_goldens``.
"""

from __future__ import annotations

import json
import os
from types import SimpleNamespace as NS

from . import ref_loader as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "chunk_cfg_goldens.json")


def load():
    R._base_stubs()
    c = os.path.join(R.REF_ROOT, "connectomics")
    R._stub("connectomics.utils", os.path.join(c, "utils"))
    R._load("connectomics.utils.channel_slices", "connectomics/utils/channel_slices.py")
    R._load("connectomics.utils.model_outputs", "connectomics/utils/model_outputs.py")
    R._stub("connectomics.data", os.path.join(c, "data"))
    R._stub("connectomics.data.processing", os.path.join(c, "data", "processing"))
    R._load("connectomics.data.processing.affinity", "connectomics/data/processing/affinity.py")
    return R._load("connectomics.inference.chunk_grid", "connectomics/inference/chunk_grid.py")


def make_cfg(case: dict):
    """the ONE place a golden case turns into a config object (the test imports it)"""
    targets = [dict(name=t) for t in case.get("extra", ())]
    if case.get("offsets"):
        targets.append(dict(name="affinity", kwargs=dict(offsets=case["offsets"], affinity_mode=case.get("mode", "deepem"))))
    return NS(data=NS(label_transform=NS(targets=targets, stack_outputs=True)),
              model=NS(out_channels=case.get("out_channels", 3), heads={}),
              inference=NS(model=NS(crop_pad=case.get("crop_pad"), select_channel=case.get("select_channel")),
                           save_backend=case.get("backend", "h5"),
                           chunking=NS(chunk_size=case.get("chunk_size", [64, 64, 64]), axes=case.get("axes", "all"),
                                       output_mode=case.get("output_mode", "decoded"))))


CASES = [
    dict(name="plain", final=[100, 200, 300]),
    dict(name="crop3", crop_pad=[1, 2, 3], final=[40, 50, 60], chunk_size=[64, 32, 128]),
    dict(name="crop6", crop_pad=[1, 2, 3, 4, 5, 6], final=[512, 512, 512], chunk_size=[128, 256, 1024], axes="z"),
    dict(name="deepem3", offsets=["1-0-0", "0-1-0", "0-0-1"], mode="deepem", final=[10, 20, 30]),
    dict(name="deepem_long", offsets=["1-0-0", "0-1-0", "0-0-1", "3-0-0", "0-9-0", "0-0-27"], mode="deepem", crop_pad=[2, 0, 1],
         out_channels=6, final=[100, 100, 100], output_mode="raw_prediction"),
    dict(name="deepem_neg", offsets=[[-1, 0, 0], [0, -2, 0], [0, 0, 4]], mode="deepem", final=[64, 64, 64]),
    dict(name="deepem_select", offsets=["1-0-0", "0-1-0", "0-0-1", "4-0-0", "0-4-0", "0-0-4"], mode="deepem", out_channels=6,
         select_channel="0:3", final=[64, 64, 64]),
    dict(name="deepem_after_binary", extra=["binary"], offsets=["2-0-0", "0-3-0"], mode="deepem", select_channel=[0, 2],
         final=[64, 64, 64]),
    dict(name="banis", offsets=["1-0-0", "0-1-0", "0-0-1"], mode="banis", crop_pad=[4, 4, 4], final=[64, 64, 64]),
]

BAD = [dict(name="bad_crop", crop_pad=[1, 2]), dict(name="bad_axes", axes="y", final=[8, 8, 8]),
       dict(name="bad_mode", output_mode="logits"), dict(name="bad_backend", backend="zarr")]


def main():
    cg = load()
    out = {"cases": [], "bad": []}
    for case in CASES:
        cfg = make_cfg(case)
        rec = dict(case=case)
        rec["normalize_crop_pad"] = cg.normalize_crop_pad(case.get("crop_pad"))
        rec["selected_offsets"] = cg.resolve_selected_affinity_offsets(cfg)
        rec["global_crop"] = cg.resolve_global_prediction_crop(cfg)
        rec["chunk_shape"] = cg.resolve_chunk_shape(cfg, case["final"])
        rec["h5_chunks"] = cg.resolve_h5_spatial_chunks(case["final"])
        rec["output_mode"] = cg.resolve_chunk_output_mode(cfg)
        cg.validate_chunked_output_format(cfg)
        out["cases"].append(rec)
    for case in BAD:
        cfg = make_cfg(case)
        calls = {"bad_crop": lambda: cg.normalize_crop_pad(case["crop_pad"]),
                 "bad_axes": lambda: cg.resolve_chunk_shape(cfg, case["final"]),
                 "bad_mode": lambda: cg.resolve_chunk_output_mode(cfg),
                 "bad_backend": lambda: cg.validate_chunked_output_format(cfg)}
        try:
            calls[case["name"]]()
            msg = None
        except ValueError as e:
            msg = str(e)
        out["bad"].append(dict(case=case, error=msg))
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {OUT}: {len(out['cases'])} cases, {len(out['bad'])} error cases")


if __name__ == "__main__":
    main()

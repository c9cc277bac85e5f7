"""Generate tests/golden/tta_goldens.npz from the REAL reference files (build container only):

    python -m oracle.make_tta_goldens

``connectomics/inference/tta_combinations.py`` and ``tta_ensemble.py`` are executed in place (``ref_loader``; their
``tta_affinity`` dependency — which pulls the data package — is replaced by a three-line stand-in for the ``ViewValidity``
container, none of its logic is on the fully-valid-channel path)."""

from __future__ import annotations

import json
import os
from types import SimpleNamespace as NS

import numpy as np
import torch

from . import ref_loader as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

COMBO_CASES = [
    dict(flip_axes="all", rotation90_axes=None),
    dict(flip_axes=None, rotation90_axes=None),
    dict(flip_axes="none", rotation90_axes="all"),
    dict(flip_axes="all", rotation90_axes="all"),
    dict(flip_axes=[[0], [1, 2]], rotation90_axes=[[1, 2]], rotate90_k=[0, 1]),
    dict(flip_axes=[[2], [2], [0, 0, 1]], rotation90_axes=[[2, 1], [1, 2]], rotate90_k=[1, 5, 3]),
    dict(flip_axes=[], rotation90_axes=[[0, 2]], rotate90_k=[2]),
]
MODE_CASES = [("mean", 3), ("max", 2), ([["0:2", "min"], ["2:", "mean"]], 3), ([[":", "max"]], 4), (["min"], 2),
              ([["0", "mean"], ["1:3", "max"], ["-1", "min"]], 4)]


class _VV:
    def __init__(self, channels):
        self.channels = tuple(channels)

    def select(self, idx):
        return self if idx is None else _VV([self.channels[i] for i in idx])


def load():
    R._base_stubs()
    R._stub("connectomics.utils", os.path.join(R.REF_ROOT, "connectomics", "utils"))
    R._load("connectomics.utils.channel_slices", "connectomics/utils/channel_slices.py")
    tc = R._load("connectomics.inference.tta_combinations", "connectomics/inference/tta_combinations.py")
    R._stub("connectomics.inference.tta_affinity", ValidityEntry=object, ViewValidity=_VV)
    te = R._load("connectomics.inference.tta_ensemble", "connectomics/inference/tta_ensemble.py")
    return tc, te


def main():
    assert R.available(), "needs /root/reference"
    tc, te = load()
    g = {}
    combos = []
    for case in COMBO_CASES:
        for sd in (3, 2):
            if sd == 2 and any(isinstance(v, list) and any(isinstance(e, list) and any(a > 1 for a in e) for e in v)
                               for v in case.values()):
                continue
            if sd == 2 and case.get("rotation90_axes") not in (None, "all", "none"):
                continue
            out = tc.resolve_tta_augmentation_combinations(NS(**case), spatial_dims=sd)
            combos.append({"cfg": case, "spatial_dims": sd,
                           "combos": [[list(f), (list(p) if p is not None else None), int(k)] for f, p, k in out]})
    g["combos_json"] = np.frombuffer(json.dumps(combos).encode(), dtype=np.uint8)
    modes = [{"mode": m, "num_channels": n, "map": tc._resolve_ensemble_mode_map(m, n)} for m, n in MODE_CASES]
    g["modes_json"] = np.frombuffer(json.dumps(modes).encode(), dtype=np.uint8)
    # streaming ensemble of the real accumulator over 5 views of a small tensor, per mode and dtype
    torch.manual_seed(0)
    views = [torch.randn(1, 3, 4, 5, 6) for _ in range(5)]
    g["ens_views"] = torch.stack(views).numpy()
    for name, dt in (("f32", torch.float32), ("f16", torch.float16), ("bf16", torch.bfloat16)):
        for mi, mode_cfg in enumerate(["mean", "min", "max", [["0:1", "max"], ["1:", "mean"]]]):
            mode_map = tc._resolve_ensemble_mode_map(mode_cfg, 3)
            acc = te.TTAEnsembleAccumulator((1, 3, 4, 5, 6), dtype=dt, device=torch.device("cpu"), mode_map=mode_map,
                                            partial_channels=[], distributed_sharding=False, max_views=5)
            for v in views:
                acc.add(v.to(dt), _VV([None, None, None]))
            g[f"ens_{name}_{mi}"] = acc.finalize().float().numpy()
    np.savez_compressed(os.path.join(OUT, "tta_goldens.npz"), **g)
    print("wrote", os.path.join(OUT, "tta_goldens.npz"), {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()

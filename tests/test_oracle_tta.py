"""TTA host logic and oracle pinned to golden vectors produced by the REAL reference files
(oracle/make_tta_goldens.py runs connectomics/inference/tta_combinations.py and tta_ensemble.py in place)."""
import json
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import tta_oracle as O
from pytorch_connectomics_b200.inference import tta as T


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "tta_goldens.npz"))


def test_augmentation_combinations_match_reference(gold):
    cases = json.loads(bytes(gold["combos_json"]).decode())
    assert len(cases) >= 9
    for case in cases:
        got = T.resolve_tta_augmentation_combinations(NS(**case["cfg"]), spatial_dims=case["spatial_dims"])
        got = [[list(f), (list(p) if p is not None else None), int(k)] for f, p, k in got]
        assert got == case["combos"], case["cfg"]
    # the Lucchi++ tutorial: 8 flip views; flips x all rotation planes: 32 unique views
    assert len(T.resolve_tta_augmentation_combinations(NS(flip_axes="all", rotation90_axes=None), spatial_dims=3)) == 8
    assert len(T.resolve_tta_augmentation_combinations(NS(flip_axes="all", rotation90_axes="all"), spatial_dims=3)) == 32


def test_ensemble_mode_map_matches_reference(gold):
    for case in json.loads(bytes(gold["modes_json"]).decode()):
        assert T._resolve_ensemble_mode_map(case["mode"], case["num_channels"]) == case["map"]
    with pytest.raises(ValueError, match="does not cover channels"):
        T._resolve_ensemble_mode_map([["0:1", "mean"]], 3)
    with pytest.raises(ValueError, match="Unknown ensemble mode"):
        T._resolve_ensemble_mode_map([[":", "median"]], 2)
    with pytest.raises(ValueError):
        T.resolve_tta_augmentation_combinations(NS(flip_axes=[[3]]), spatial_dims=3)
    with pytest.raises(ValueError, match="exactly 2 axes"):
        T.resolve_tta_augmentation_combinations(NS(rotation90_axes=[[0, 1, 2]]), spatial_dims=3)


def test_channel_selectors():
    assert T.resolve_channel_range(":", num_channels=4) == (0, 4)
    assert T.resolve_channel_range("1:-1", num_channels=4) == (1, 3)
    assert T.resolve_channel_range(-1, num_channels=4) == (3, 4)
    assert T.resolve_channel_indices([0, "2", -1], num_channels=4) == [0, 2, 3]
    assert T.resolve_channel_indices(None, num_channels=4) == [0, 1, 2, 3]     # utils/channel_slices.py:207-208: None = every channel
    for bad in ("4:", "2:1", "1:2:3", ""):
        with pytest.raises(ValueError):
            T.resolve_channel_range(bad, num_channels=4)
    codes, scales = T.resolve_activation_codes([{"channels": "0:2", "activation": "sigmoid"},
                                                {"channels": 2, "activation": "scale_sigmoid:0.5"},
                                                {"channels": [3], "activation": "tanh"}], 5)
    assert codes == [1, 1, 2, 3, 0] and scales[2] == 0.5
    codes, _scales, groups = T.resolve_activation_specs([{"channels": ":", "activation": "softmax"}], 3)
    assert codes == [4, 4, 4] and groups[0] == [0, 1, 2]
    with pytest.raises(ValueError, match="Unknown activation"):
        T.resolve_activation_codes([{"channels": ":", "activation": "relu"}], 3)


@pytest.mark.parametrize("name,dt", [("f32", torch.float32), ("f16", torch.float16), ("bf16", torch.bfloat16)])
def test_oracle_fold_matches_reference_accumulator(gold, name, dt):
    views = torch.from_numpy(gold["ens_views"])
    for mi, mode_cfg in enumerate(["mean", "min", "max", [["0:1", "max"], ["1:", "mean"]]]):
        modes = T._resolve_ensemble_mode_map(mode_cfg, 3)
        acc = None
        for n_prev in range(views.shape[0]):
            acc = O.fold(acc, views[n_prev].to(dt), modes, n_prev)
        assert torch.equal(acc.float(), torch.from_numpy(gold[f"ens_{name}_{mi}"])), (name, mode_cfg)


def test_oracle_view_roundtrip():
    x = torch.arange(2 * 3 * 4 * 5 * 6, dtype=torch.float32).reshape(2, 3, 4, 5, 6)
    for f, p, k in T.resolve_tta_augmentation_combinations(NS(flip_axes="all", rotation90_axes="all"), spatial_dims=3):
        assert torch.equal(O.invert_view(O.view(x, f, p, k), f, p, k), x)

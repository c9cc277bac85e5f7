"""GPU parity of the TTA kernels (csrc/tta_kernels.cu) against the CPU oracle and the reference-generated goldens:
views are pure gathers (bit-exact), the fold without activations is bit-exact in fp32 / fp16 / bf16 (same rounding
points as torch's elementwise kernels); activations use expf/tanhf (tolerance 2e-6 relative in fp32, 1 ulp in half)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import tta_oracle as O
from pytorch_connectomics_b200.inference import tta as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
def test_views_match_torch(dt):
    torch.manual_seed(0)
    x = torch.randn(2, 3, 5, 6, 7).to(dt)
    xg = x.to(DEV)
    for f, p, k in T.resolve_tta_augmentation_combinations(NS(flip_axes="all", rotation90_axes="all"), spatial_dims=3):
        got = T.apply_view(xg, f, p, k)
        want = O.view(x, f, p, k)
        assert got.shape == want.shape and torch.equal(got.cpu(), want), (f, p, k)
    got = T.apply_view(xg, [1], (2, 1), 3)        # reversed plane order
    assert torch.equal(got.cpu(), O.view(x, [1], (2, 1), 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.apply_view(x, [0], None, 0)


@pytest.mark.parametrize("name,dt", [("f32", torch.float32), ("f16", torch.float16), ("bf16", torch.bfloat16)])
def test_fold_matches_reference_accumulator_goldens(name, dt):
    gold = np.load(os.path.join(GOLDEN, "tta_goldens.npz"))
    views = torch.from_numpy(gold["ens_views"]).to(DEV)
    for mi, mode_cfg in enumerate(["mean", "min", "max", [["0:1", "max"], ["1:", "mean"]]]):
        it = iter(range(views.shape[0]))
        ens = T.TTAEnsemble(NS(enabled=True, flip_axes=[[0]] * 4, ensemble_mode=mode_cfg), output_dtype=dt)
        # identity "network" that returns pre-inverted views: feed view(v_i) so that the fold un-views it back to v_i
        combos = ens.combinations(5)
        assert len(combos) == 5

        def net(x_aug, combos=combos, it=it):
            i = next(it)
            f, p, k = combos[i]
            return T.apply_view(views[i].to(dt), f, p, k) if (f or p) else views[i].to(dt)

        got = ens.predict(views[0].to(dt), net)
        assert got.dtype == dt
        assert torch.equal(got.float().cpu(), torch.from_numpy(gold[f"ens_{name}_{mi}"])), (name, mode_cfg)


def _net(x):   # 1 -> 3 channels, position dependent so that wrong index maps cannot cancel
    d, h, w = x.shape[2:]
    ramp = (torch.arange(d, device=x.device, dtype=torch.float32).view(1, 1, d, 1, 1) * 0.01
            + torch.arange(h, device=x.device, dtype=torch.float32).view(1, 1, 1, h, 1) * 0.02
            + torch.arange(w, device=x.device, dtype=torch.float32).view(1, 1, 1, 1, w) * 0.03).to(x.dtype)
    return torch.cat([x * 2.0 - 1.0 + ramp, x * x + ramp, -x + 0.5 * ramp], dim=1)


@pytest.mark.parametrize("dt,cfg,acts,select", [
    (torch.float32, dict(flip_axes="all", ensemble_mode="mean"), None, None),
    (torch.float32, dict(flip_axes="all", rotation90_axes=[[1, 2]], ensemble_mode=[["0:2", "min"], ["2:", "max"]]), None, None),
    (torch.float32, dict(flip_axes=[[0], [1, 2]], rotation90_axes="all", rotate90_k=[1, 3], ensemble_mode="mean"),
     [{"channels": "0:2", "activation": "sigmoid"}, {"channels": 2, "activation": "scale_sigmoid:0.5"}], [2, 0]),
    (torch.float16, dict(flip_axes="all", ensemble_mode="mean"), [{"channels": ":", "activation": "tanh"}], "1:"),
    (torch.bfloat16, dict(flip_axes="all", rotation90_axes=[[0, 1]], ensemble_mode="max"), None, None),
])
def test_tta_ensemble_matches_oracle(dt, cfg, acts, select):
    torch.manual_seed(1)
    x = torch.rand(1, 1, 6, 6, 8).to(dt)       # rotations in a non-square plane change the view shape
    tta_cfg = NS(enabled=True, **cfg)
    ens = T.TTAEnsemble(tta_cfg, channel_activations=acts, select_channel=select, output_dtype=dt)
    got = ens.predict(x.to(DEV), _net)
    combos = T.resolve_tta_augmentation_combinations(tta_cfg, spatial_dims=3)
    codes, scales = T.resolve_activation_codes(acts, 3)
    sel = T.resolve_channel_indices(select, num_channels=3, context="select")
    nsel = 3 if sel is None else len(sel)
    modes = T._resolve_ensemble_mode_map(cfg["ensemble_mode"], nsel)
    want = O.tta_predict(x, _net, combos, modes, codes, scales, sel, dt)
    assert got.shape == want.shape and got.dtype == dt
    if acts is None and dt == torch.float32:
        assert torch.equal(got.cpu(), want)
    else:
        tol = {torch.float32: 2e-6, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}[dt]
        assert torch.allclose(got.float().cpu(), want.float(), rtol=tol, atol=tol)


def test_tta_over_sliding_window_engine():
    from pytorch_connectomics_b200.inference.window import EagerSlidingWindowEngine
    torch.manual_seed(2)
    x = torch.rand(1, 1, 20, 24, 28)
    eng_kw = dict(roi_size=(16, 16, 16), sw_batch_size=3, overlap=0.5, mode="bump", padding_mode="constant", cval=0.0)
    eng = EagerSlidingWindowEngine(**eng_kw)
    tta_cfg = NS(enabled=True, flip_axes="all", ensemble_mode="mean")
    acts = [{"channels": ":", "activation": "sigmoid"}]
    got = T.TTAEnsemble(tta_cfg, channel_activations=acts).predict(x.to(DEV), lambda v: eng(inputs=v, network=_net))
    from oracle import window_oracle as WO
    combos = T.resolve_tta_augmentation_combinations(tta_cfg, spatial_dims=3)
    want = O.tta_predict(x, lambda v: WO.eager_sliding_window(v, _net, sw_batch_size=3, **{k: eng_kw[k] for k in ("roi_size", "overlap", "mode", "padding_mode", "cval")}),
                         combos, ["mean"] * 3, [1, 1, 1], [1.0] * 3, None, torch.float32)
    assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5)

"""``TTAPredictor(cfg, sliding_inferer, forward_fn)`` (reference ``inference/tta.py:67-79,1619-1666``): the cfg-driven object
the reference's inference loop calls.  CPU part: input normalisation, mask validation / alignment / application (pure index and
pointwise logic, restated from ``tta.py:465-601,1568-1617`` — known answers written out here), switches.  GPU part: ``predict``
against ``TTAEnsemble`` (which the golden-pinned tests of ``test_tta_gpu.py`` / ``test_tta_affinity.py`` cover) and against a
hand-composed mask application."""

from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from pytorch_connectomics_b200.inference.tta import TTAPredictor


def _cfg(tta=None, acts=None, select=None, odt=None, **model):
    return NS(model=NS(**{"out_channels": 3, "primary_head": None, **model}),
              inference=NS(test_time_augmentation=tta, sliding_window=NS(keep_input_on_cpu=False),
                           model=NS(channel_activations=acts, select_channel=select, output_dtype=odt, head=None)))


def test_input_normalisation_and_switches():
    p = TTAPredictor(_cfg(NS(enabled=True, patch_first_local=True, distributed_sharding=True)), None, lambda x: x)
    assert p._normalize_input(torch.zeros(4, 5, 6)).shape == (1, 1, 4, 5, 6)
    assert p._normalize_input(torch.zeros(2, 4, 5, 6)).shape == (2, 1, 4, 5, 6)
    assert p._normalize_input(torch.zeros(2, 3, 4, 5, 6)).shape == (2, 3, 4, 5, 6)
    with pytest.raises(ValueError, match="3D, 4D, or 5D"):
        p._normalize_input(torch.zeros(4, 5))
    assert not p._is_patch_first_local_tta_enabled()              # needs a sliding inferer
    assert TTAPredictor(p.cfg, object(), lambda x: x)._is_patch_first_local_tta_enabled()
    assert not p.is_distributed_sharding_enabled()                # no process group
    assert not p.should_skip_postprocess_on_rank()
    assert TTAPredictor(NS(), None, None)._get_tta_cfg() is None
    with pytest.raises(RuntimeError, match="keep_input_on_cpu"):
        cfg = _cfg()
        cfg.inference.sliding_window.keep_input_on_cpu = True
        TTAPredictor(cfg, None, lambda x: x)._run_network(torch.zeros(1, 1, 2, 2, 2))


def test_head_selection_and_activation_types():
    acts = [dict(channels="0:2", activation="sigmoid"), dict(channels="2:3", activation="tanh")]
    out = {"output": {"aff": torch.zeros(1, 3, 2, 2, 2), "sdt": torch.ones(1, 1, 2, 2, 2)}}
    p = TTAPredictor(_cfg(acts=acts, select=[2, 0], primary_head="aff"), None, lambda x: out)
    got = p._sliding_window_predict(torch.zeros(1, 1, 2, 2, 2))
    assert got is out["output"]["aff"] and p.channel_activation_types == ["tanh", "sigmoid"]     # selection order
    r = TTAPredictor(_cfg(primary_head="aff"), None, lambda x: out)
    r._requested_output_head_override = "sdt"
    assert r._sliding_window_predict(torch.zeros(1, 1, 2, 2, 2)) is out["output"]["sdt"]
    r._requested_output_head_override = "nope"
    with pytest.raises(ValueError, match="requested_head"):
        r._sliding_window_predict(torch.zeros(1, 1, 2, 2, 2))
    q = TTAPredictor(_cfg(), None, lambda x: x)
    q._sliding_window_predict(torch.zeros(1, 3, 2, 2, 2))
    assert q.channel_activation_types is None


def test_mask_validation_alignment_and_application():
    p = TTAPredictor(_cfg(NS(enabled=False, apply_mask=True)), None, lambda x: x)
    pred = torch.arange(2 * 3 * 2 * 4 * 4, dtype=torch.float32).reshape(2, 3, 2, 4, 4) + 1
    m3 = torch.zeros(2, 4, 4)
    m3[:, 1:3, 1:3] = 7.0                                           # any positive value counts
    m = p._validate_and_prepare_mask(m3, pred)
    assert m.shape == (2, 1, 2, 4, 4) and m.dtype == pred.dtype and set(m.unique().tolist()) == {0.0, 1.0}
    assert p._validate_and_prepare_mask([[m3.numpy()]], pred).shape == (2, 1, 2, 4, 4)          # collated containers unwrap
    assert p._validate_and_prepare_mask(torch.ones(2, 2, 4, 4), pred).shape == (2, 1, 2, 4, 4)  # (B, D, H, W)
    with pytest.raises(ValueError, match="Mask is None"):
        p._validate_and_prepare_mask(None, pred)
    with pytest.raises(ValueError, match="rank"):
        p._validate_and_prepare_mask(torch.ones(4, 4), pred)
    with pytest.raises(ValueError, match="Mask batch 3"):
        p._validate_and_prepare_mask(torch.ones(3, 1, 2, 4, 4), pred)
    with pytest.raises(ValueError, match="Mask channels 2"):
        p._validate_and_prepare_mask(torch.ones(2, 2, 2, 4, 4), pred)
    with pytest.raises(ValueError, match="exactly match"):
        p._validate_and_prepare_mask(torch.ones(1, 1, 2, 4, 5), pred)
    # align_to_image: centre crop where the mask is larger (6 -> 4: drop one voxel per side), zero pad where it is smaller
    # (3 -> 4: the odd voxel goes behind)
    big = torch.zeros(1, 1, 2, 6, 3)
    big[..., 1:5, :] = 1.0
    al = p._validate_and_prepare_mask(big, pred, align_to_image=True)
    assert al.shape == (2, 1, 2, 4, 4) and al[0, 0, 0].tolist() == [[1, 1, 1, 0]] * 4
    # application: plain product without activation metadata; tanh channels are filled with -1 outside the mask
    out = p._apply_mask_to_result(pred.clone(), m3, False)
    assert torch.equal(out, pred * m)
    p.channel_activation_types = ["sigmoid", None, "tanh"]
    out = p._apply_mask_to_result(pred.clone(), m3, False)
    assert torch.equal(out[:, :2], (pred * m)[:, :2]) and torch.equal(out[:, 2:], pred[:, 2:] * m + (1 - m) * -1.0)
    per_channel = torch.stack([m3, 1 - m3.clamp(max=1), m3], 0)[None].expand(2, 3, 2, 4, 4)
    out = p._apply_mask_to_result(pred.clone(), per_channel, False)
    assert torch.equal(out[:, 1], pred[:, 1] * (per_channel[:, 1] > 0)) and float(out[0, 2, 0, 0, 0]) == -1.0
    assert p._apply_mask_to_result(pred, "not a mask", False) is pred                            # TypeError -> skipped
    off = TTAPredictor(_cfg(NS(enabled=False, apply_mask=False)), None, lambda x: x)
    assert off._apply_mask_to_result(pred, m3, False) is pred and p._apply_mask_to_result(pred, None, False) is pred


def test_native_plan_passthrough_only_for_known_pure_forwards():
    """The engine may take the in-library tile loop only when ``forward_fn`` IS the pcb200 module, its bound ``forward``, or the
    reference LightningModule's pass-through ``forward``; a requested head or any other wrapper keeps the generic loop."""
    class Plan:
        head_channels = [3]

    class Net(torch.nn.Module):
        def native_plan(self):
            return Plan()

        def forward(self, x):
            return x

    class ConnectomicsModule(torch.nn.Module):           # name-matched stand-in of training/lightning/model.py
        def __init__(self):
            super().__init__()
            self.model = Net()

        def forward(self, x):
            return self.model(x)

    class Other(ConnectomicsModule):
        pass

    acts = [dict(channels="2:3", activation="tanh")]
    for fn, expect in ((Net(), True), (Net().forward, True), (ConnectomicsModule().forward, True), (Other().forward, False),
                       (lambda x: x, False), (Net().native_plan, False)):
        p = TTAPredictor(_cfg(acts=acts), object(), fn)
        plan = p._engine_network.native_plan()
        assert (plan is not None) == expect
        if expect:                                       # the activation metadata the mask step needs is known without a call
            assert p.channel_activation_types == [None, None, "tanh"]
    p = TTAPredictor(_cfg(), object(), Net())
    p._requested_output_head_override = "aff"
    assert p._engine_network.native_plan() is None

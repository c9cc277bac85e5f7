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


# ----------------------------------------------------------------------------- GPU: predict == TTAEnsemble + mask
def _net(t):
    return torch.cat([t * 0.5 + 0.25, 1.0 - t, t * t], 1)


@pytest.mark.gpu
def test_predict_matches_ensemble_and_applies_mask():
    from pytorch_connectomics_b200.inference import window as W
    from pytorch_connectomics_b200.inference.tta import TTAEnsemble
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    x = torch.rand(1, 1, 16, 16, 16, device=dev)
    acts = [dict(channels="0:2", activation="sigmoid"), dict(channels="2:3", activation="tanh")]
    tta = NS(enabled=True, flip_axes="all", rotation90_axes=None, rotate90_k=None, ensemble_mode="mean", apply_mask=True,
             patch_first_local=False, distributed_sharding=False)
    cfg = _cfg(tta, acts=acts, select=[2, 0])
    p = TTAPredictor(cfg, None, _net)
    got = p.predict(x[0, 0])                                         # (D, H, W) input is expanded
    want = TTAEnsemble(tta, channel_activations=acts, select_channel=[2, 0], output_dtype=torch.float32).predict(x, _net)
    assert got.shape == (1, 2, 16, 16, 16) and torch.equal(got, want)
    assert p.channel_activation_types == ["tanh", "sigmoid"]
    mask = (torch.rand(16, 16, 16, device=dev) > 0.5).float()
    masked = p.predict(x, mask=mask)
    m = mask[None, None]
    assert torch.equal(masked[:, 0:1], want[:, 0:1] * m + (1 - m) * -1.0) and torch.equal(masked[:, 1:2], want[:, 1:2] * m)
    # TTA disabled: one identity view == apply_preprocessing(network(x)); fp16 output dtype from the config
    cfg_off = _cfg(NS(enabled=False), acts=acts, odt="float16")
    q = TTAPredictor(cfg_off, None, _net)
    plain = q.predict(x)
    raw = _net(x)
    ref = torch.cat([torch.sigmoid(raw[:, :2]), torch.tanh(raw[:, 2:])], 1)
    assert plain.dtype == torch.float16 and torch.allclose(plain.float(), ref, atol=2e-3)
    assert torch.equal(q.apply_preprocessing(raw), plain)
    # through a sliding-window engine, volume-first and patch-first (flip views commute with a pointwise network, so both
    # equal the plain ensemble up to blending round-off)
    eng = W.EagerSlidingWindowEngine(roi_size=(16, 16, 16), sw_batch_size=2, overlap=0.5, mode="constant",
                                     padding_mode="constant", cval=0.0)
    big = torch.rand(1, 1, 24, 16, 32, device=dev)
    direct = TTAPredictor(cfg, None, _net).predict(big)
    vol_first = TTAPredictor(cfg, eng, _net).predict(big)
    tta.patch_first_local = True
    patch_first = TTAPredictor(cfg, eng, _net).predict(big)
    assert torch.allclose(vol_first, direct, atol=1e-5) and torch.allclose(patch_first, direct, atol=1e-5)


@pytest.mark.gpu
def test_lazy_seam_runs_patches_through_the_predictor(tmp_path):
    """lazy.py:1038,1187-1194: with TTA / activations / channel selection configured, the lazy engine sends every patch batch
    through ``TTAPredictor(cfg, None, forward_fn).predict`` (views + activations + selection + mask per PATCH) before
    blending.  With a pointwise forward and flip views every view of a patch gives the same values, so the result is
    activation(net(volume))[selected] * mask up to blending round-off."""
    from pytorch_connectomics_b200.inference import lazy as Z
    dev = "cuda:0"
    vol = np.random.RandomState(5).rand(24, 16, 40).astype(np.float32)
    mask = (np.random.RandomState(6).rand(24, 16, 40) > 0.3).astype(np.float32)
    np.save(tmp_path / "v.npy", vol)
    np.save(tmp_path / "m.npy", mask)
    sw = NS(window_size=[16, 16, 16], overlap=0.5, blending="constant", sw_batch_size=2, padding_mode="constant", cval=0.0,
            snap_to_edge=False, target_context=[], border_mask=None, distributed_sharding=False)
    acts = [dict(channels="0:2", activation="sigmoid"), dict(channels="2:3", activation="tanh")]
    cfg = NS(model=NS(output_size=[16, 16, 16], arch=NS(type="mednext"), primary_head=None),
             data=NS(dataloader=NS(batch_size=1, patch_size=[16, 16, 16]), data_transform=NS()),
             inference=NS(sliding_window=sw, model=NS(output_dtype=None, channel_activations=acts, select_channel=[2, 0], head=None),
                          test_time_augmentation=NS(enabled=True, flip_axes="all", rotation90_axes=None, rotate90_k=None,
                                                    ensemble_mode="mean", apply_mask=True)))
    got = Z.lazy_predict_volume(cfg, _net, str(tmp_path / "v.npy"), mask_path=str(tmp_path / "m.npy"), device=dev)
    x = torch.from_numpy(vol)[None, None]
    raw = _net(x)
    m = torch.from_numpy(mask)[None, None]
    want = torch.cat([torch.tanh(raw[:, 2:3]) * m + (1 - m) * -1.0, torch.sigmoid(raw[:, 0:1]) * m], 1)
    assert got.shape == want.shape and torch.allclose(got, want, atol=2e-5)

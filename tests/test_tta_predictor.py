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


# ----------------------------------------------------------------------------- predict() glue on the CPU, oracle ensemble
from cpu_doubles import oracle_ensemble_predict as _oracle_ensemble_predict  # noqa: E402


def test_predict_glue_against_the_oracle_chain(monkeypatch):
    """``TTAPredictor.predict`` end to end on the CPU with the fold kernels replaced by the oracle chain: what is checked is
    everything the predictor adds — config nodes -> ensemble arguments, input normalisation, routing through a sliding
    inferer object, named heads, the disabled-TTA path, mask + tanh fill, output dtype."""
    from oracle import tta_oracle as O
    from pytorch_connectomics_b200.inference import tta as T
    monkeypatch.setattr(T.TTAEnsemble, "predict", _oracle_ensemble_predict)
    net = O.ramp_network(3)
    torch.manual_seed(2)
    x = torch.rand(1, 1, 6, 8, 8)
    acts = [dict(channels="0:2", activation="sigmoid"), dict(channels="2:3", activation="tanh")]
    tta = NS(enabled=True, flip_axes="all", rotation90_axes=[[1, 2]], rotate90_k=None,
             ensemble_mode=[["0:1", "min"], ["1:2", "mean"]], apply_mask=True, patch_first_local=False, distributed_sharding=False)
    cfg = _cfg(tta, acts=acts, select=[2, 0])
    combos = T.resolve_tta_augmentation_combinations(tta, spatial_dims=3)
    assert len(combos) > 8
    want = O.tta_predict(x, net, combos, ["min", "mean"], [1, 1, 3], [1.0, 1.0, 1.0], [2, 0], torch.float32)
    calls = []

    def engine(inputs, network):                     # a sliding inferer as the predictor sees it: keyword call, callable network
        calls.append(tuple(inputs.shape))
        assert network.native_plan() is None         # a plain function has no native plan: generic loop
        return network(inputs)

    for inferer in (None, engine):
        p = TTAPredictor(cfg, inferer, net)
        got = p.predict(x[0, 0])                     # (D, H, W) in
        assert got.shape == (1, 2, 6, 8, 8) and torch.equal(got, want)
        assert p.channel_activation_types == ["tanh", "sigmoid"]
    assert len(calls) == len(combos)                 # one engine call per view
    mask = (torch.rand(6, 8, 8) > 0.4).float()
    m = mask[None, None]
    masked = TTAPredictor(cfg, None, net).predict(x, mask=[[mask.numpy()]])          # collated container around the mask
    assert torch.equal(masked[:, :1], want[:, :1] * m + (1 - m) * -1.0) and torch.equal(masked[:, 1:], want[:, 1:] * m)
    tta.apply_mask = False
    assert torch.equal(TTAPredictor(cfg, None, net).predict(x, mask=mask), want)
    # TTA disabled: one pass = apply_preprocessing; fp16 output dtype from inference.model.output_dtype
    off = _cfg(NS(enabled=False), acts=acts, odt="float16")
    plain = TTAPredictor(off, None, net).predict(x)
    ref = O.preprocess_specs(net(x), [([0, 1], "sigmoid"), ([2], "tanh")], None, torch.float16)
    assert plain.dtype == torch.float16 and torch.equal(plain, ref)
    # named heads: {"output": {head: tensor}}; requested_head overrides model.primary_head for one call only
    heads = lambda t: {"output": {"aff": net(t), "sdt": net(t)[:, :1] * 2.0}}
    ph = TTAPredictor(_cfg(NS(enabled=False), primary_head="aff"), None, heads)
    assert torch.equal(ph.predict(x), net(x)) and torch.equal(ph.predict(x, requested_head="sdt"), net(x)[:, :1] * 2.0)
    assert ph._requested_output_head_override is None
    with pytest.raises(ValueError, match="requested_head"):
        ph.predict(x, requested_head="nope")


def test_lazy_seam_through_the_predictor_on_cpu_doubles(tmp_path, monkeypatch):
    """The cfg-driven lazy engine (lazy.py:986-1258) with TTA, activations, channel selection and a mask file configured:
    every patch batch goes through ``TTAPredictor(cfg, None, forward_fn).predict`` (lazy.py:1038,1187-1194).  Kernel-calling
    helpers are replaced by the oracle stand-ins (``tests/cpu_doubles.py``), so this is the CPU run of the host logic whose
    GPU run is ``test_zz_first_run_gpu.py::test_lazy_seam_runs_patches_through_the_predictor``.  Pointwise forward + flip
    views: every view of a patch gives the same values, so the blended result is activation(net(volume))[selected] * mask."""
    import cpu_doubles
    from pytorch_connectomics_b200.inference import lazy as Z
    cpu_doubles.install(monkeypatch)
    vol = np.random.RandomState(5).rand(12, 8, 20).astype(np.float32)
    mask = (np.random.RandomState(6).rand(12, 8, 20) > 0.3).astype(np.float32)
    np.save(tmp_path / "v.npy", vol)
    np.save(tmp_path / "m.npy", mask)
    sw = NS(window_size=[8, 8, 8], overlap=0.5, blending="constant", sw_batch_size=2, padding_mode="constant", cval=0.0,
            snap_to_edge=False, target_context=[], border_mask=None, distributed_sharding=False)
    acts = [dict(channels="0:2", activation="sigmoid"), dict(channels="2:3", activation="tanh")]
    cfg = NS(model=NS(output_size=[8, 8, 8], arch=NS(type="mednext"), primary_head=None),
             data=NS(dataloader=NS(batch_size=1, patch_size=[8, 8, 8]), data_transform=NS()),
             inference=NS(sliding_window=sw, model=NS(output_dtype=None, channel_activations=acts, select_channel=[2, 0], head=None),
                          test_time_augmentation=NS(enabled=True, flip_axes="all", rotation90_axes=None, rotate90_k=None,
                                                    ensemble_mode="mean", apply_mask=True)))
    net = lambda t: torch.cat([t * 0.5 + 0.25, 1.0 - t, t * t], 1)
    got = Z.lazy_predict_volume(cfg, net, str(tmp_path / "v.npy"), mask_path=str(tmp_path / "m.npy"), device="cpu")
    x = torch.from_numpy(vol)[None, None]
    raw = net(x)
    m = torch.from_numpy(mask)[None, None]
    want = torch.cat([torch.tanh(raw[:, 2:3]) * m + (1 - m) * -1.0, torch.sigmoid(raw[:, 0:1]) * m], 1)
    assert got.shape == want.shape and torch.allclose(got, want, atol=2e-6)
    # without TTA / activations / selection the predictor is not built and the direct route gives net(volume) * mask
    cfg.inference.test_time_augmentation.enabled = False
    cfg.inference.model.channel_activations, cfg.inference.model.select_channel = None, None
    plain = Z.lazy_predict_volume(cfg, net, str(tmp_path / "v.npy"), mask_path=str(tmp_path / "m.npy"), device="cpu")
    assert torch.allclose(plain, raw * m, atol=2e-6)


# ----------------------------------------------------------------------------- against the REAL TTAPredictor (tta.py)
def _real_tta():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference is only present in the build container")
    return ref_loader.ref_tta()


_ACTS = [dict(channels="0:2", activation="sigmoid"), dict(channels="2:3", activation="tanh")]
_CASES = {
    "flips_mean": dict(tta=dict(enabled=True, flip_axes="all", rotation90_axes=None, rotate90_k=None, ensemble_mode="mean"), acts=_ACTS),
    "rot_modes_select": dict(tta=dict(enabled=True, flip_axes=[[0], [1, 2]], rotation90_axes=[[1, 2]], rotate90_k=[0, 1, 3],
                                      ensemble_mode=[["0:1", "min"], ["1:2", "max"]]), acts=_ACTS, select=[2, 0]),
    "softmax_scale": dict(tta=dict(enabled=True, flip_axes=[[2]], rotation90_axes=None, rotate90_k=None, ensemble_mode="max"),
                          acts=[dict(channels=[0, 1], activation="softmax"), dict(channels="2", activation="scale_sigmoid:0.5")], select="1:"),
    "single_view": dict(tta=dict(enabled=True, flip_axes=None, rotation90_axes=None, rotate90_k=None, ensemble_mode="mean"), acts=_ACTS),
    "disabled_fp16": dict(tta=dict(enabled=False), acts=_ACTS, odt="float16"),
    "no_activations": dict(tta=dict(enabled=True, flip_axes=[[1]], rotation90_axes=None, rotate90_k=None, ensemble_mode="mean"), acts=None),
}


@pytest.mark.parametrize("case", sorted(_CASES))
@pytest.mark.parametrize("with_mask", [False, True])
def test_predictor_equals_the_real_tta_predictor(case, with_mask, monkeypatch):
    """This package's `TTAPredictor` (fold kernels replaced by the oracle chain, `tests/cpu_doubles.py`) against the REAL
    `connectomics/inference/tta.py::TTAPredictor` executed in place, same config object, same forward: equal predictions
    (views, activations incl. softmax / scale_sigmoid, channel selection, per-channel ensemble modes, output dtype), equal
    `channel_activation_types`, equal masked result incl. the tanh fill."""
    from oracle import tta_oracle as O
    from pytorch_connectomics_b200.inference import tta as T
    R = _real_tta()
    monkeypatch.setattr(T.TTAEnsemble, "predict", _oracle_ensemble_predict)
    spec = _CASES[case]
    cfg = _cfg(NS(apply_mask=True, patch_first_local=False, distributed_sharding=False, **spec["tta"]), acts=spec.get("acts"),
               select=spec.get("select"), odt=spec.get("odt"))
    net = O.ramp_network(3)
    rs = np.random.RandomState(3)
    x = torch.from_numpy(rs.rand(1, 1, 6, 8, 8).astype(np.float32))
    mask = torch.from_numpy((rs.rand(6, 8, 8) > 0.4).astype(np.float32)) if with_mask else None
    want = R.TTAPredictor(cfg, None, net).predict(x.clone(), mask=mask)
    mine = TTAPredictor(cfg, None, net)
    got = mine.predict(x.clone(), mask=mask)
    assert got.shape == want.shape and got.dtype == want.dtype
    tol = 2e-3 if want.dtype == torch.float16 else 2e-6
    assert torch.allclose(got.float(), want.float(), rtol=tol, atol=tol), float((got.float() - want.float()).abs().max())
    ref_types = R.TTAPredictor(cfg, None, net)
    ref_types.predict(x.clone())
    norm = lambda ts: None if ts is None else [None if t is None else t.split(":")[0] for t in ts]
    assert norm(mine.channel_activation_types) == norm(ref_types.channel_activation_types)


def test_predictor_with_the_real_sliding_engine_and_named_heads(monkeypatch):
    """Volume-first TTA through a sliding-window engine and named-head selection, both predictors driving the REAL
    `EagerSlidingWindowEngine` of `window.py` on the CPU."""
    from oracle import ref_loader, tta_oracle as O
    from pytorch_connectomics_b200.inference import tta as T
    R = _real_tta()
    W = ref_loader.ref_window()
    monkeypatch.setattr(T.TTAEnsemble, "predict", _oracle_ensemble_predict)
    tta = NS(enabled=True, flip_axes="all", rotation90_axes=None, rotate90_k=None, ensemble_mode="mean", apply_mask=True,
             patch_first_local=False, distributed_sharding=False)
    cfg = _cfg(tta, acts=_ACTS, select=[2, 0])
    cfg.model.heads = {"aff": NS(out_channels=3), "sdt": NS(out_channels=1)}
    cfg.model.primary_head = "aff"
    net = O.ramp_network(3)
    heads = lambda t: {"output": {"aff": net(t), "sdt": net(t)[:, :1] * 2.0}}
    eng_kw = dict(roi_size=(4, 8, 8), sw_batch_size=2, overlap=0.5, mode="constant", padding_mode="constant", cval=0.0,
                  sw_device=None, output_device=None)
    x = torch.from_numpy(np.random.RandomState(4).rand(1, 1, 10, 8, 8).astype(np.float32))
    want = R.TTAPredictor(cfg, W.EagerSlidingWindowEngine(**eng_kw), heads).predict(x.clone())
    got = TTAPredictor(cfg, W.EagerSlidingWindowEngine(**eng_kw), heads).predict(x.clone())
    assert torch.allclose(got, want, rtol=2e-6, atol=2e-6)
    # an explicit head for one call: activations are written for the 3-channel head, so ask for it by name
    want = R.TTAPredictor(cfg, None, heads).predict(x.clone(), requested_head="aff")
    got = TTAPredictor(cfg, None, heads).predict(x.clone(), requested_head="aff")
    assert torch.allclose(got, want, rtol=2e-6, atol=2e-6)

    def outcome(fn):
        try:
            return ("ok", tuple(fn().shape))
        except (ValueError, TypeError) as e:
            return (type(e).__name__, str(e))

    for bad in ("nope", "", "aff,sdt"):
        assert outcome(lambda: R.TTAPredictor(cfg, None, heads).predict(x.clone(), requested_head=bad)) == \
            outcome(lambda: TTAPredictor(cfg, None, heads).predict(x.clone(), requested_head=bad)), bad


@pytest.mark.parametrize("case", ["volume_first", "direct_rot", "disabled"])
def test_predictor_reproduces_the_real_predictor_goldens_on_cpu(case, monkeypatch):
    """`tests/golden/tta_predictor_goldens.npz` (REAL `TTAPredictor` + REAL `EagerSlidingWindowEngine`, bump blending, mask,
    `oracle/make_tta_predictor_goldens.py`).  This package's predictor with the oracle ensemble, driving the real engine on
    the CPU, reproduces the volume-first / direct / disabled goldens; the patch-first golden needs the kernels (GPU test)."""
    import os
    from conftest import GOLDEN
    from oracle import make_tta_predictor_goldens as G, ref_loader, tta_oracle as O
    from pytorch_connectomics_b200.inference import tta as T
    if not ref_loader.available():
        pytest.skip("the real window.py engine is only present in the build container")
    monkeypatch.setattr(T.TTAEnsemble, "predict", _oracle_ensemble_predict)
    gold = np.load(os.path.join(GOLDEN, "tta_predictor_goldens.npz"))
    spec = G.CASES[case]
    x, mask = G.inputs()
    engine = ref_loader.ref_window().EagerSlidingWindowEngine(sw_device=None, output_device=None, **G.ENGINE) if spec["engine"] else None
    got = TTAPredictor(G.make_cfg(spec), engine, O.ramp_network(3)).predict(x, mask=mask if spec["mask"] else None)
    assert torch.allclose(got, torch.from_numpy(gold[case]), rtol=2e-6, atol=2e-6)

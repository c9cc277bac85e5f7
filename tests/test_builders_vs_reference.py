"""The architecture-builder seam against the REAL ``connectomics/models/architectures/mednext_models.py`` (builders, wrappers,
task heads, validation), executed in place over a stand-in ``nnunet_mednext`` module whose network classes are the oracle
restatements (``oracle/ref_loader.py::ref_mednext_models``; the third-party package cannot be installed offline).  For every
configuration: same wrapper type and attributes, same ``get_model_info()``, same ``state_dict`` keys and shapes (so reference
checkpoints load into this package's modules and vice versa), and — for bad configurations — the same error.  CPU only: the
modules are built, not run (``tests/test_mednext_gpu.py`` runs them)."""

from types import SimpleNamespace as NS

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is only present in the build container")


def _cfg(arch="mednext", heads=None, primary=None, ds=False, out_channels=2, in_channels=1, **mednext):
    return NS(model=NS(arch=NS(type=arch), in_channels=in_channels, out_channels=out_channels, heads=heads, primary_head=primary,
                       mednext=NS(**mednext), loss=NS(deep_supervision=ds)))


def _outcome(fn):
    try:
        return ("ok", fn())
    except (ValueError, TypeError, KeyError, NotImplementedError) as e:
        return (type(e).__name__, str(e))


def _describe(m):
    sd = m.state_dict()
    d = dict(type=type(m).__name__, ds=m.supports_deep_supervision, scales=m.output_scales, info=m.get_model_info(),
             keys=[(k, tuple(v.shape)) for k, v in sd.items()], repr_head=repr(m).split("(")[0])
    for attr in ("primary_head", "feature_channels"):
        if hasattr(m, attr):
            d[attr] = getattr(m, attr)
    if hasattr(m, "head_specs"):
        d["head_specs"] = {k: dict(v) if isinstance(v, dict) else v for k, v in m.head_specs.items()}
    if hasattr(m, "heads"):
        d["heads"] = sorted(m.heads.keys())
    return d


HEADS = {"aff": NS(out_channels=3, num_blocks=1, hidden_channels=16), "sdt": dict(out_channels=1, num_blocks=0)}

GOOD = [
    _cfg(size="S", kernel_size=3), _cfg(size="B", kernel_size=5, ds=True), _cfg(size="S", kernel_size=3, ds=True, out_channels=3),
    _cfg(size="S", kernel_size=3, checkpoint_style="outside_block", in_channels=2),
    _cfg(size="S", kernel_size=3, heads=HEADS, primary="aff"), _cfg(size="S", kernel_size=3, heads={"only": NS(out_channels=2)}),
    _cfg("mednext_custom", base_channels=16, exp_r=2, kernel_size=3, block_counts=[1] * 9),
    _cfg("mednext_custom", base_channels=16, exp_r=[2, 3, 4, 4, 4, 4, 4, 3, 2], kernel_size=5, block_counts=[1, 2, 1, 1, 1, 1, 1, 2, 1], ds=True,
         do_res=False, do_res_up_down=False),
    _cfg("mednext_custom", base_channels=16, exp_r=2, kernel_size=3, block_counts=[1] * 9, norm="layer", checkpoint_style="outside_block"),
    _cfg("mednext_custom", base_channels=16, exp_r=2, kernel_size=3, block_counts=[1] * 9, heads=HEADS, primary="sdt"),
    # model.mednext.dim / grn (mednext_models.py:461-462): Conv2d parameter shapes, grn_gamma / grn_beta keys, heads that inherit both
    _cfg("mednext_custom", base_channels=16, exp_r=2, kernel_size=3, block_counts=[1] * 9, dim="2d", ds=True),
    _cfg("mednext_custom", base_channels=16, exp_r=2, kernel_size=5, block_counts=[1] * 9, grn=True, norm="layer"),
    _cfg("mednext_custom", base_channels=16, exp_r=2, kernel_size=3, block_counts=[1] * 9, dim="2d", grn=True, heads=HEADS, primary="aff"),
]

BAD = [
    _cfg(size="XL", kernel_size=3), _cfg(size="S", kernel_size=4), _cfg(size="S", kernel_size=3, checkpoint_style="inside"),
    _cfg("mednext_custom", base_channels=16, dim="4d"), _cfg("mednext_custom", base_channels=16, norm="batch"),
    _cfg("mednext_custom", base_channels=16, block_counts=[1] * 8),
    _cfg(size="S", kernel_size=3, heads=HEADS, primary="aff", ds=True),                       # heads reject deep-supervision trunks
    _cfg(size="S", kernel_size=3, heads={"a": NS(out_channels=0)}), _cfg(size="S", kernel_size=3, heads={"a": NS(out_channels=2, num_blocks=-1)}),
    _cfg(size="S", kernel_size=3, heads={"a": NS(out_channels=2, hidden_channels=64)}),
    _cfg(size="S", kernel_size=3, heads=HEADS, primary="nope"),
]


def _builders():
    import pytorch_connectomics_b200.architectures as A
    R = ref_loader.ref_mednext_models()
    return {"mednext": (R.build_mednext, A.get_architecture_builder("mednext")),
            "mednext_custom": (R.build_mednext_custom, A.get_architecture_builder("mednext_custom"))}


@pytest.mark.parametrize("i", range(len(GOOD)))
def test_built_modules_match_the_real_builders(i):
    cfg = GOOD[i]
    real_fn, ours_fn = _builders()[cfg.model.arch.type]
    torch.manual_seed(0)
    real = real_fn(cfg)
    ours = ours_fn(cfg)
    want, got = _describe(real), _describe(ours)
    assert want == got, {k: (want[k], got[k]) for k in want if want[k] != got.get(k)}
    ours.load_state_dict(real.state_dict(), strict=True)          # a reference checkpoint loads, and the other way round
    real.load_state_dict(ours.state_dict(), strict=True)
    assert getattr(real.model, "outside_block_checkpointing", False) == getattr(ours.model, "outside_block_checkpointing", False)


@pytest.mark.parametrize("i", range(len(BAD)))
def test_bad_configurations_fail_like_the_real_builders(i):
    cfg = BAD[i]
    real_fn, ours_fn = _builders()[cfg.model.arch.type]
    want, got = _outcome(lambda: type(real_fn(cfg)).__name__), _outcome(lambda: type(ours_fn(cfg)).__name__)
    assert want[0] != "ok", "the reference accepts this configuration; move it to GOOD"
    assert want == got


# ----------------------------------------------------------------------------- monai_unet (BASELINE configs[0])
def _ucfg(filters=(16, 32, 64), input_size=(32, 64, 64), in_channels=1, out_channels=1, **monai):
    return NS(model=NS(arch=NS(type="monai_unet"), in_channels=in_channels, out_channels=out_channels,
                       input_size=list(input_size) if input_size else None, monai=NS(filters=list(filters), **monai)))


UNET_GOOD = [_ucfg(num_res_units=1, kernel_size=3, norm="batch", dropout=0.0),          # tutorials/minimal.yaml
             _ucfg(filters=(8, 16, 32, 64), in_channels=2, out_channels=3),              # defaults: 2 residual units
             _ucfg(filters=(16, 32), num_res_units=0), _ucfg(input_size=None, spatial_dims=3, num_res_units=1),
             _ucfg(num_res_units=1, norm="instance", dropout=0.1),                        # no norm parameters; dropout module kept
             _ucfg(num_res_units=1, norm="group", num_groups=4, out_channels=4), _ucfg(num_res_units=2, norm="group", out_channels=8),
             # UpsampleModeUNet (monai_models.py:84-139): the REAL _get_up_layer override runs over the oracle's UpSample
             _ucfg(num_res_units=1, upsample_mode="nontrainable"), _ucfg(filters=(8, 16, 32, 64), num_res_units=0, upsample_mode="nontrainable",
                                                                         upsample_interp_mode="nearest", upsample_align_corners=None),
             _ucfg(num_res_units=2, upsample_mode="nontrainable", upsample_interp_mode="trilinear", upsample_align_corners=False, out_channels=3)]


@pytest.mark.parametrize("i", range(len(UNET_GOOD)))
def test_monai_unet_matches_the_real_builder(i):
    """`build_monai_unet` (monai_models.py:197-250) executed in place over the oracle `UNet` / `ResidualUnit`: wrapper type,
    model info, MONAI `state_dict` keys and shapes, checkpoints load both ways."""
    import pytorch_connectomics_b200.architectures as A
    R = ref_loader.ref_monai_models()
    cfg = UNET_GOOD[i]
    real, ours = R.build_monai_unet(cfg), A.get_architecture_builder("monai_unet")(cfg)
    want, got = _describe(real), _describe(ours)
    assert want == got, {k: (want[k], got[k]) for k in want if want[k] != got.get(k)}
    ours.load_state_dict(real.state_dict(), strict=True)
    real.load_state_dict(ours.state_dict(), strict=True)


def test_oracle_upsample_mode_unet_equals_the_real_class():
    """the oracle's restated `UpsampleModeUNet` (what the GPU tests compare with, where the reference is absent) against the
    REAL class executed in place: same `state_dict`, same output for the same weights, in both modes"""
    import torch
    from oracle import monai_unet_oracle as UO
    R = ref_loader.ref_monai_models()
    for mode, interp, align in (("nontrainable", "linear", True), ("nontrainable", "nearest", None), ("deconv", "linear", True)):
        kw = dict(spatial_dims=3, in_channels=1, out_channels=2, channels=[8, 16, 32], strides=[2, 2], num_res_units=1,
                  kernel_size=3, norm="batch", dropout=0.0, upsample_mode=mode, upsample_interp_mode=interp,
                  upsample_align_corners=align)
        torch.manual_seed(0)
        real, ours = R.UpsampleModeUNet(**kw).eval(), UO.UpsampleModeUNet(**kw).eval()
        ours.load_state_dict(real.state_dict(), strict=True)
        x = torch.rand(1, 1, 16, 16, 16)
        with torch.no_grad():
            assert torch.equal(real(x), ours(x)), mode


def test_multihead_golden_is_reproducible_from_the_real_wrapper():
    """`tests/golden/multihead_golden.npz` (`oracle/make_multihead_goldens.py`): rebuilt here from the REAL
    `MedNeXtMultiHeadWrapper` over the oracle trunk with the name-seeded weights — the committed fixture is what the generator
    produces, bit for bit on the same platform (tolerance 1e-6 relative for BLAS differences)."""
    import os
    import numpy as np
    from conftest import GOLDEN
    from oracle import make_multihead_goldens as G
    M = ref_loader.ref_mednext_models()
    net = M.build_mednext_custom(G.make_cfg()).train()
    G.fill_deterministic(net)
    x, g = G.inputs()
    out = net(x)["output"]
    gold = np.load(os.path.join(GOLDEN, "multihead_golden.npz"))
    for k in G.HEADS:
        assert np.allclose(out[k].detach().numpy(), gold[f"out_{k}"], rtol=1e-5, atol=1e-6)
    assert list(gold["grad_names"]) == [n for n, p in net.named_parameters() if not n.startswith("model.out_") and n != "model.dummy_tensor"]


def test_group_norm_needs_divisible_channels_like_the_real_builder():
    """`norm: group` with the default 8 groups and a 1-channel output: MONAI's top up-sampling layer normalises the OUTPUT
    channels, so torch refuses the GroupNorm — same refusal from both builders."""
    import pytorch_connectomics_b200.architectures as A
    R = ref_loader.ref_monai_models()
    cfg = _ucfg(num_res_units=1, norm="group", out_channels=1)
    want = _outcome(lambda: R.build_monai_unet(cfg))
    got = _outcome(lambda: A.get_architecture_builder("monai_unet")(cfg))
    assert want[0] == "ValueError" and want == got

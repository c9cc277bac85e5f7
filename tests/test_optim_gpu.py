"""Fused AdamW + clip + EMA kernel (csrc/optim_kernels.cu) against torch.optim.AdamW + clip_grad_norm_ + the reference's
EMA update (training/lightning/callbacks.py:869-907) on the CPU, same parameter groups (training/optimization/build.py:69-111).
fp32 arithmetic, torch's op order: tolerance 2e-6 relative (documented in the kernel: (1 - lr*wd) and lr/bc1 are formed
in fp32 instead of Python doubles)."""
from types import SimpleNamespace as NS

import pytest
import torch

from pytorch_connectomics_b200.training import FlatGradArena, FusedAdamW, build_fused_adamw, reference_param_groups

DEV = "cuda:0"


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv3d(2, 5, 3), torch.nn.GroupNorm(5, 5), torch.nn.Conv3d(5, 3, 1, bias=False),
                               torch.nn.Linear(7, 3))


def test_reference_param_groups_follow_build_optimizer():
    m = _model()
    groups = reference_param_groups(m, 1e-3, 0.01, weight_decay_norm=0.0, weight_decay_bias=0.002, bias_lr_factor=2.0)
    by = {id(g["params"][0]): g for g in groups}
    assert by[id(m[0].weight)]["weight_decay"] == 0.01 and by[id(m[0].bias)]["weight_decay"] == 0.002
    assert by[id(m[0].bias)]["lr"] == 2e-3 and by[id(m[1].weight)]["weight_decay"] == 0.0 and by[id(m[1].bias)]["weight_decay"] == 0.0
    assert len(groups) == len(list(m.parameters()))


@pytest.mark.gpu
@pytest.mark.parametrize("clip,ema,world", [(0.0, None, 1), (0.5, 0.99, 1), (1.0, 0.999, 4)])
def test_fused_adamw_matches_torch(clip, ema, world):
    ref = _model()
    net = _model().to(DEV)
    lr, wd = 1e-2, 0.05
    kw = dict(weight_decay_norm=0.0, weight_decay_bias=0.01, bias_lr_factor=2.0)
    topt = torch.optim.AdamW(reference_param_groups(ref, lr, wd, **kw), lr=lr, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    arena = FlatGradArena(net.parameters())
    groups = reference_param_groups(net, lr, wd, **kw)
    fopt = FusedAdamW(groups, betas=(0.9, 0.99), eps=1e-8, max_grad_norm=clip, ema_decay=ema, arena=arena, world_size=world)
    ema_ref = {k: v.detach().clone() for k, v in ref.named_parameters()} if ema is not None else None
    g = torch.Generator().manual_seed(1)
    for it in range(4):
        grads = [torch.randn(p.shape, generator=g) * (3.0 if it == 2 else 0.3) for p in ref.parameters()]
        for p, gr in zip(ref.parameters(), grads):
            p.grad = gr.clone()
        if clip > 0:
            torch.nn.utils.clip_grad_norm_(ref.parameters(), clip)
        topt.step()
        if ema_ref is not None:
            for k, p in ref.named_parameters():
                ema_ref[k].mul_(ema).add_(p.detach(), alpha=1.0 - ema)
        arena.zero()
        for p, gr in zip(net.parameters(), grads):       # the all-reduced SUM over `world` identical ranks
            p.grad.copy_((gr * world).to(DEV))
        fopt.step(grads_are_summed=True)
    torch.cuda.synchronize()
    for (k, a), b in zip(ref.named_parameters(), net.parameters()):
        assert torch.allclose(b.cpu(), a.detach(), rtol=2e-6, atol=2e-7), (k, (b.cpu() - a).abs().max())
    if ema is not None:
        views = fopt.ema_tensors()
        for (k, _), b in zip(ref.named_parameters(), net.parameters()):
            assert torch.allclose(views[id(b)].cpu(), ema_ref[k], rtol=2e-6, atol=2e-7), k
        before = [p.detach().clone() for p in net.parameters()]
        fopt.swap_ema()
        for (k, _), b in zip(ref.named_parameters(), net.parameters()):
            assert torch.allclose(b.cpu(), ema_ref[k], rtol=2e-6, atol=2e-7)
        fopt.swap_ema()
        assert all(torch.equal(a, b.detach()) for a, b in zip(before, net.parameters()))
    assert float(fopt.step_count) == 4.0
    # the modules still see the flat storage: a forward pass works and parameters are views of ONE arena
    assert net(torch.zeros(1, 2, 5, 5, 9, device=DEV)).shape == (1, 3, 3, 3, 3)
    assert all(fopt.flat.data_ptr() <= p.data_ptr() < fopt.flat.data_ptr() + 4 * fopt.n and p.data_ptr() % 16 == 0
               for p in net.parameters())


@pytest.mark.gpu
def test_build_fused_adamw_from_cfg_and_graph_capture():
    net = _model().to(DEV)
    cfg = NS(optimization=NS(optimizer=NS(name="adamw", lr=1e-3, weight_decay=0.01, betas=[0.9, 0.999], eps=1e-8),
                             gradient_clip_val=1.0))
    arena = FlatGradArena(net.parameters())
    opt = build_fused_adamw(cfg, net, arena=arena, ema_decay=0.99)
    assert opt.max_grad_norm == 1.0 and float(opt.seg_wd[2]) == 0.0        # GroupNorm weight: weight_decay_norm default 0
    arena.buffer.normal_()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        opt.step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        opt.step()
    p0 = opt.flat.clone()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    assert float(opt.step_count) == 3.0 and not torch.equal(p0, opt.flat)
    with pytest.raises(NotImplementedError):
        build_fused_adamw(NS(optimization=NS(optimizer=NS(name="sgd"))), net)


@pytest.mark.gpu
def test_fused_adamw_trains_mednext_like_torch_adamw():
    """Whole loop on the engine: two identical tiny MedNeXts, one stepped by torch.optim.AdamW (same param groups), one by
    the fused kernel — after 3 steps the weights agree and the fused model's forward uses the UPDATED weights (the
    kernel-layout weight cache is invalidated by the parameter epoch).  Both optimizers are fed the SAME gradients (the
    engine's backward of the fused model, copied into the torch-stepped twin): Adam's first steps move every weight by
    +-lr whatever the gradient's size, so two separately differentiated replicas diverge by O(lr) wherever a
    gradient element is rounding noise around zero (measured 1e-4 on conv1.weight) — that would test the sign of noise, not
    the optimizer."""
    from pytorch_connectomics_b200.architectures import mednext as PM

    def make():
        torch.manual_seed(3)
        return PM.MedNeXt(1, 16, 1, exp_r=2, kernel_size=3, deep_supervision=False, do_res=True, do_res_up_down=True,
                          block_counts=[1] * 9).to(DEV).train()

    a, b = make(), make()
    lr, wd = 1e-3, 0.01
    topt = torch.optim.AdamW(reference_param_groups(a, lr, wd), lr=lr, weight_decay=wd)
    arena = FlatGradArena(b.parameters())
    fopt = FusedAdamW(reference_param_groups(b, lr, wd), arena=arena)
    assert all(p.data_ptr() % 16 == 0 for p in b.parameters())      # dummy_tensor (1 element) must not misalign the rest
    torch.manual_seed(4)
    xs = [torch.rand(1, 1, 32, 32, 32, device=DEV).half() for _ in range(3)]
    ts = [(torch.rand(1, 1, 32, 32, 32, device=DEV) > 0.8).float() for _ in range(3)]
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    with torch.no_grad():
        out0 = b(xs[0]).clone()
    for x, t in zip(xs, ts):
        fopt.zero_grad()
        bce(b(x).float(), t).backward()
        for pa, pb in zip(a.parameters(), b.parameters()):
            # a slice of the arena that stayed exactly zero is a parameter autograd never reached (dummy_tensor): the
            # reference leaves its .grad None and torch.optim skips it — so does the fused kernel (seg_active)
            unused = pb.grad is None or not bool(pb.grad.any())
            pa.grad = None if unused else pb.grad.detach().clone()
        topt.step()
        fopt.step()
    torch.cuda.synchronize()
    for (k, pa), pb in zip(a.named_parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=0, atol=2e-5), (k, (pa - pb).abs().max())
    with torch.no_grad():
        out_a, out_b = a(xs[0]), b(xs[0])
    assert not torch.equal(out_b, out0)                      # the forward sees the stepped weights
    # weights agree to 2e-5; the two forwards still round to bf16 at different places (|out| ~ 3: one bf16 ulp = 1.6e-2)
    rel = float((out_a.float() - out_b.float()).norm() / out_a.float().norm())
    assert rel < 1e-2, rel


@pytest.mark.gpu
def test_arena_train_step_accumulates_like_one_large_batch():
    """trainer.py:314-334 `accumulate_grad_batches`: two micro-batches of 2 with the loss divided by 2 move the weights exactly
    like one batch of 4 (mean loss; GroupNorm is per sample, so the per-sample arithmetic is identical — only the fp32 order of
    the gradient sums differs), the optimizer runs once per window, and the grad-norm clip sees the accumulated gradient."""
    from pytorch_connectomics_b200.architectures import mednext as PM
    from pytorch_connectomics_b200.training import ArenaTrainStep

    def make():
        torch.manual_seed(5)
        net = PM.MedNeXt(1, 16, 1, exp_r=2, kernel_size=3, deep_supervision=False, do_res=True, do_res_up_down=True,
                         block_counts=[1] * 9).to(DEV).train()
        arena = FlatGradArena(net.parameters())
        opt = FusedAdamW(reference_param_groups(net, 1e-3, 0.01), arena=arena, max_grad_norm=1.0)
        return net, opt

    bce = torch.nn.functional.binary_cross_entropy_with_logits
    loss_fn = lambda out, t: bce(out.float(), t)
    torch.manual_seed(6)
    x = torch.rand(4, 1, 32, 32, 32, device=DEV).half()
    t = (torch.rand(4, 1, 32, 32, 32, device=DEV) > 0.8).float()
    a, oa = make()
    b, ob = make()
    one = ArenaTrainStep(a, loss_fn, oa)
    acc = ArenaTrainStep(b, loss_fn, ob, accumulate_grad_batches=2)
    la = one(x, t)
    assert not acc.will_step
    l0 = acc(x[:2], t[:2])
    assert acc.optimizer_steps == 0 and acc.will_step
    l1 = acc(x[2:], t[2:])
    assert acc.optimizer_steps == 1 and one.optimizer_steps == 1
    torch.cuda.synchronize()
    assert abs(float(la) - 0.5 * (float(l0) + float(l1))) < 1e-5
    assert abs(float(oa.grad_norm()) - float(ob.grad_norm())) < 1e-4 * float(oa.grad_norm())
    for (k, pa), pb in zip(a.named_parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=0, atol=2e-6), (k, float((pa - pb).abs().max()))
    assert acc.flush() is None and acc(x[:2], t[:2]) is not None and acc.flush() == 1 and acc.optimizer_steps == 2
    with pytest.raises(ValueError):
        ArenaTrainStep(a, loss_fn, oa, accumulate_grad_batches=0)

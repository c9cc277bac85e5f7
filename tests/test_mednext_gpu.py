"""GPU parity: MedNeXt kernels (C ABI) vs the CPU oracle on the same seeded inputs/weights.

Tolerance.  north_star asks for 1e-3 relative vs the reference's own PyTorch path.  The engine
computes in bf16 (config 2 is bf16; tcgen05 kind::f16 with fp32 accumulation), as does the
reference under `precision: bf16-mixed` — and bf16 storage alone rounds every activation to 2^-9
= 2e-3 relative.  So each check states BOTH numbers: the engine's relative L2 error against the
fp32 oracle, and the same error for the oracle run under torch.autocast(bfloat16) (= the reference's
own bf16 path); the engine must not be further from fp32 than 1.5x the reference's bf16 path
(+1e-3 absolute slack).

That bound alone is loose (it admits an engine error of ~1.7e-2 and hides a bug confined to a few rows inside an
L2 norm), so every block / network check ALSO compares against the ROUNDING-MATCHED oracle
(`oracle.mednext_oracle.bf16_matched()`: the same stock torch ops with bf16 rounding exactly where the engine stores
bf16 and bf16 pointwise weights, fp32 in between): rel-L2 <= MATCHED_REL (2e-3 per block) AND max-abs <= MATCHED_ULPS
bf16 ulps of the output range (a localised error cannot hide in a max).  Through a whole network the bound is 8e-3:
two bf16 pipelines that differ only in accumulation order still round ~10 % of the elements of every stored tensor to
the neighbouring bf16 value (measured per block: ~1.2e-3 rel-L2), and 21 blocks of MedNeXt-S compound that as sqrt(21)
(measured on the B200: 5.5e-3 at 160^3, 6.2e-3 at 32^3, max-abs 3.3 ulps of range)."""
import os

import numpy as np
import pytest
import torch

from oracle import mednext_oracle as OM
from pytorch_connectomics_b200 import _lib as L
from pytorch_connectomics_b200.architectures import _mednext_ops as ops
from pytorch_connectomics_b200.architectures import mednext as PM

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def cl(x):   # NCDHW fp32 cpu -> channels-last bf16 cuda
    return ops.as_channels_last(x.to(DEV))


def ncdhw(x):
    return x.permute(0, 4, 1, 2, 3).float().cpu()


def autocast_ref(mod, *a):
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        return mod(*a).float()


MATCHED_REL, MATCHED_REL_NET, MATCHED_ULPS = 2e-3, 8e-3, 8.0


def matched_ref(mod, *a, **kw):
    with torch.no_grad(), OM.bf16_matched():
        out = mod(*a, **kw)
    return [t.float() for t in out] if isinstance(out, (list, tuple)) else out.float()


def check_matched(got, want_m, label, rel_bound=MATCHED_REL, ulps=MATCHED_ULPS):
    g, w = got.float().cpu(), want_m.float().cpu()
    e = float((g - w).norm() / w.norm().clamp_min(1e-12))
    rng = float(w.abs().max())
    mx = float((g - w).abs().max())
    n_ulp = mx / (rng * 2.0 ** -8) if rng > 0 else 0.0
    print(f"{label}: vs rounding-matched oracle rel-L2 {e:.3e} (bound {rel_bound:.0e}), max-abs {mx:.3e} = {n_ulp:.2f} "
          f"bf16 ulps of range {rng:.3g} (bound {ulps})")
    assert e <= rel_bound, (label, "matched rel-L2", e)
    assert n_ulp <= ulps, (label, "matched max-abs ulps", n_ulp)
    return e


def check(got, want32, want_bf16, label, slack=1e-3, want_m=None, rel_bound=MATCHED_REL):
    e, e_ref = rel(got, want32), rel(want_bf16, want32)
    print(f"{label}: engine rel-L2 {e:.3e}   reference-bf16-path rel-L2 {e_ref:.3e}")
    assert e <= 1.5 * e_ref + slack, (label, e, e_ref)
    if want_m is not None:
        check_matched(got, want_m, label, rel_bound)
    return e


def test_device_ok():
    assert L.lib().pcb_device_ok() == 1


@pytest.mark.parametrize("cin,c", [(1, 32), (3, 16)])
def test_stem(cin, c):
    torch.manual_seed(0)
    conv = torch.nn.Conv3d(cin, c, 1)
    x = torch.rand(2, cin, 8, 12, 16)
    with torch.no_grad():
        want = conv(x)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        got = ops.stem_forward(x.to(DEV, dt), conv.weight.to(DEV), conv.bias.to(DEV))
        assert got.shape == (2, 8, 12, 16, c) and got.dtype == torch.bfloat16
        assert rel(ncdhw(got), want) < (4e-3 if dt == torch.float32 else 8e-3)


@pytest.mark.parametrize("c,k,size", [(32, 3, (10, 12, 14)), (16, 5, (9, 8, 11)), (64, 7, (8, 8, 8)), (32, 3, (7, 9, 13)),
                                      (64, 3, (11, 18, 41))])
@pytest.mark.parametrize("mode", ["same", "down", "up"])
def test_dwconv_and_stats(c, k, size, mode):
    torch.manual_seed(1)
    if mode == "same":
        conv = torch.nn.Conv3d(c, c, k, 1, k // 2, groups=c)
    elif mode == "down":
        conv = torch.nn.Conv3d(c, c, k, 2, k // 2, groups=c)
    else:
        conv = torch.nn.ConvTranspose3d(c, c, k, 2, k // 2, groups=c)
    x = torch.randn(2, c, *size)
    xq = x.bfloat16().float()
    with torch.no_grad():
        want = conv(xq)
    import ctypes
    xc = cl(x)
    w = ops.packed(conv.weight.to(DEV), "dw")
    b = conv.bias.detach().to(DEV).float().contiguous()
    y = torch.empty((2, *want.shape[2:], c), device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros((2, 2, c), device=DEV, dtype=torch.float64)
    m = {"same": L.DW_SAME, "down": L.DW_DOWN, "up": L.DW_UP}[mode]
    L.check(L.lib().pcb_dwconv_fwd(L.ptr(xc), L.ptr(w), L.ptr(b), L.ptr(y), L.ptr(stats), ctypes.c_int64(2),
                                   L.i64x(size), ctypes.c_int64(c), k, m, L.stream_ptr()), "dw")
    torch.cuda.synchronize()
    got = ncdhw(y)
    assert got.shape == want.shape
    assert rel(got, want) < 4e-3            # bf16 output rounding only (fp32 accumulate)
    yq = y.float()
    s = yq.sum(dim=(1, 2, 3)).double().cpu()
    q = (yq * yq).sum(dim=(1, 2, 3)).double().cpu()
    assert torch.allclose(stats[:, 0].cpu(), s, rtol=1e-4, atol=1e-3)
    assert torch.allclose(stats[:, 1].cpu(), q, rtol=1e-4, atol=1e-3)


def _mk_pair(kind, cin, cout, r, k):
    torch.manual_seed(2)
    cls_o = {"same": OM.MedNeXtBlock, "down": OM.MedNeXtDownBlock, "up": OM.MedNeXtUpBlock}[kind]
    cls_p = {"same": PM.MedNeXtBlock, "down": PM.MedNeXtDownBlock, "up": PM.MedNeXtUpBlock}[kind]
    o = cls_o(cin, cout, r, k, do_res=True, norm_type="group").eval()
    with torch.no_grad():   # non-trivial affine so the GroupNorm fold is exercised
        o.norm.weight.uniform_(0.5, 1.5)
        o.norm.bias.uniform_(-0.5, 0.5)
    p = cls_p(cin, cout, r, k, do_res=True, norm_type="group").eval()
    p.load_state_dict(o.state_dict(), strict=True)
    return o, p.to(DEV)


@pytest.mark.parametrize("kind,cin,cout,r,k,size", [
    ("same", 32, 32, 2, 3, (16, 16, 16)),      # level-0 shape: K=32, H=64
    ("same", 16, 16, 4, 3, (9, 10, 11)),       # ragged tile (990 voxels), K=16
    ("same", 64, 64, 3, 5, (8, 8, 8)),         # B/M exp ratios, k=5
    ("same", 256, 256, 2, 3, (4, 6, 4)),       # multi K-chunk, multi hidden-chunk
    ("same", 512, 512, 2, 3, (4, 4, 4)),       # bottleneck: Co split over 2 CTAs
    ("same", 128, 128, 8, 3, (4, 4, 6)),       # L: r=8 -> hidden 1024
    ("down", 32, 64, 2, 3, (16, 16, 16)),
    ("down", 16, 32, 4, 3, (10, 12, 14)),
    ("down", 256, 512, 2, 3, (4, 4, 4)),
    ("up", 64, 32, 2, 3, (8, 8, 8)),
    ("up", 32, 16, 4, 3, (5, 6, 7)),
    ("up", 512, 256, 2, 3, (2, 2, 2)),
    ("same", 128, 128, 2, 3, (10, 12, 14)),    # deep path: 14 ragged row tiles x column tiles
    ("down", 128, 256, 2, 3, (10, 12, 14)),    # deep path + strided res-conv K segment
    ("up", 128, 64, 2, 3, (6, 6, 6)),          # deep path: padded rows, sparse res-conv rows, BN = 64
    ("same", 256, 256, 4, 3, (6, 8, 10)),      # deep path: H = 1024, 4 column tiles, K = 256
])
def test_block_forward(kind, cin, cout, r, k, size):
    o, p = _mk_pair(kind, cin, cout, r, k)
    torch.manual_seed(3)
    x = torch.randn(2, cin, *size)
    xq = x.bfloat16().float()
    with torch.no_grad():
        want = o(xq)
        got = p(cl(x))
    want_bf = autocast_ref(o, xq)
    torch.cuda.synchronize()
    assert ncdhw(got).shape == want.shape
    check(ncdhw(got), want, want_bf, f"{kind} block C={cin}->{cout} r={r} k={k}", want_m=matched_ref(o, xq))


def test_up_block_with_fused_skip():
    o, p = _mk_pair("up", 64, 32, 2, 3)
    torch.manual_seed(4)
    x, skip = torch.randn(1, 64, 6, 6, 6), torch.randn(1, 32, 12, 12, 12)
    xq, sq = x.bfloat16().float(), skip.bfloat16().float()
    with torch.no_grad():
        want = sq + o(xq)
        got = p(cl(x), cl(skip))
    want_bf = sq + autocast_ref(o, xq)
    check(ncdhw(got), want, want_bf, "up block + skip", want_m=matched_ref(o, xq, skip=sq))
    # o = 0 planes carry the skip only (F.pad front zero, blocks.py MedNeXtUpBlock.forward)
    assert torch.equal(ncdhw(got)[:, :, 0], sq[:, :, 0])
    assert torch.equal(ncdhw(got)[:, :, :, :, 0], sq[:, :, :, :, 0])


@pytest.mark.parametrize("c,ncls", [(32, 1), (32, 3), (16, 12)])
def test_head(c, ncls):
    torch.manual_seed(5)
    o = OM.OutBlock(c, ncls)
    x = torch.randn(2, c, 6, 7, 8)
    xq = x.bfloat16().float()
    with torch.no_grad():
        want = o(xq)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        got = ops.head_forward(cl(x), o.conv_out.weight.to(DEV), o.conv_out.bias.to(DEV), dt)
        assert got.dtype == dt and got.shape == want.shape
        assert rel(got, want) < {torch.float32: 1e-5, torch.float16: 1e-3, torch.bfloat16: 6e-3}[dt]


def test_tiny_network_vs_golden_and_oracle(mednext_tiny_golden):
    g = mednext_tiny_golden
    kw = dict(in_channels=1, n_channels=16, n_classes=2, exp_r=2, kernel_size=3, deep_supervision=True,
              do_res=True, do_res_up_down=True, block_counts=[1] * 9)
    torch.manual_seed(0)
    o = OM.MedNeXt(**kw).eval()
    p = PM.MedNeXt(**kw).eval()
    p.load_state_dict(o.state_dict(), strict=True)
    p.to(DEV)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        outs = p(x.to(DEV))
        want = o(x)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        want_bf = [t.float() for t in o(x)]
    want_m = matched_ref(o, x)
    assert len(outs) == 5
    for i in range(5):
        assert outs[i].shape == want[i].shape and outs[i].dtype == torch.float32
        np.testing.assert_allclose(want[i].numpy(), g[f"out{i}"], rtol=1e-4, atol=1e-5)
        # deep-supervision heads read 16^3 .. 2^3-voxel levels of this 32^3 input: GroupNorm statistics over a few dozen
        # voxels amplify single bf16 rounding flips (the autocast path itself is 1.9e-2 off there) -> looser matched bound
        check(outs[i], want[i], want_bf[i], f"tiny MedNeXt out{i}", slack=2e-3, want_m=want_m[i],
              rel_bound=MATCHED_REL_NET if i == 0 else 3e-2)
    # forward_output(forward_features(x)) == model(x)  (reference tests/unit/test_mednext_features.py:26-39)
    with torch.no_grad():
        f = p.forward_features(x.to(DEV))
        assert f.shape == (1, 16, 32, 32, 32)
        assert torch.allclose(p.forward_output(f), outs[0], rtol=1e-5, atol=1e-6)


def test_mednext_s_shapes_and_parity():
    torch.manual_seed(0)
    o = OM.create_mednext_v1(1, 3, "S", 3, False).eval()
    p = PM.create_mednext_v1(1, 3, "S", 3, False).eval()
    p.load_state_dict(o.state_dict(), strict=True)
    p.to(DEV)
    x = torch.rand(1, 1, 32, 32, 32)
    with torch.no_grad():
        got = p(x.to(DEV).half())
        want = o(x.half().float())
    want_bf = autocast_ref(o, x.half().float())
    assert got.dtype == torch.float16 and got.shape == (1, 3, 32, 32, 32)
    check(got, want, want_bf, "MedNeXt-S 32^3", slack=2e-3, want_m=matched_ref(o, x.half().float()), rel_bound=MATCHED_REL_NET)
    with pytest.raises(ValueError):
        p(torch.rand(1, 1, 20, 32, 32, device=DEV))


def test_full_size_config2_forward_parity():
    """BASELINE configs[1] at full size: MedNeXt-S, 1x1x160^3.  One oracle forward on the host cores (tens of
    seconds) against the engine; same two-number tolerance statement as the small cases."""
    import os
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    torch.manual_seed(0)
    o = OM.create_mednext_v1(1, 1, "S", 3, False).eval()
    p = PM.create_mednext_v1(1, 1, "S", 3, False).eval()
    p.load_state_dict(o.state_dict(), strict=True)
    p.to(DEV)
    x = torch.rand(1, 1, 160, 160, 160).half().float()
    with torch.no_grad():
        got = p(x.to(DEV).half())
        want = o(x)
        f1 = p.forward_features(x.to(DEV).half())
        f2 = p.forward_features(x.to(DEV).half())
    want_bf = autocast_ref(o, x)
    assert got.shape == (1, 1, 160, 160, 160) and torch.isfinite(got).all()
    check(got, want, want_bf, "MedNeXt-S 160^3 (config 2)", slack=2e-3, want_m=matched_ref(o, x), rel_bound=MATCHED_REL_NET)
    # run-to-run determinism of the whole trunk at full size (fp64 statistics, ordered blending of nothing else)
    assert (f1.float() - f2.float()).abs().max().item() <= 1e-2 * f1.float().abs().max().item()


def _blob_volume(n, side, seed):
    """learnable synthetic segmentation task: smooth random field -> image = field + noise, label = field > median"""
    g = torch.Generator().manual_seed(seed)
    f = torch.rand(n, 1, side, side, side, generator=g)
    for _ in range(3):
        f = torch.nn.functional.avg_pool3d(f, 5, 1, 2, count_include_pad=False)
    f = (f - f.mean()) / f.std()
    img = (f * 0.25 + 0.5 + 0.05 * torch.randn(f.shape, generator=g)).clamp(0, 1)
    return img, (f > 0).float()


def test_jaccard_replay_after_training():
    """north_star: "matching reference Jaccard on Lucchi++ synthetic replay".  The HF checkpoint is not available
    offline, so the replay is: train MedNeXt-S with THIS engine for 200 AdamW steps on a learnable synthetic task (logits
    separate, unlike random-init weights whose logits sit at 0), load the trained weights into the fp32 CPU oracle
    (strict), predict a held-out volume with both, threshold the sigmoid at 0.5 (evaluation/metric_execution.py:178-196)
    and compare the Jaccard against the labels: |J_engine - J_oracle| <= 1e-3, and the two masks agree (J >= 0.995)."""
    import os
    from oracle.window_oracle import binary_jaccard
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    torch.manual_seed(0)
    p = PM.create_mednext_v1(1, 1, "S", 3, False).to(DEV).train()
    opt = torch.optim.AdamW(p.parameters(), lr=1e-3, weight_decay=0.01)
    imgs, labs = _blob_volume(8, 48, seed=1)
    imgs, labs = imgs.to(DEV), labs.to(DEV)
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    first = last = None
    for it in range(200):
        i = (2 * it) % 8
        opt.zero_grad(set_to_none=True)
        loss = bce(p(imgs[i:i + 2].half()).float(), labs[i:i + 2])
        loss.backward()
        opt.step()
        if it == 0:
            first = float(loss)
        last = float(loss)
    print(f"training loss {first:.4f} -> {last:.4f}")
    assert last < 0.5 * first, (first, last)          # the task is learnable and the engine's gradients train it
    p.eval()
    o = OM.create_mednext_v1(1, 1, "S", 3, False).eval()
    o.load_state_dict({k: v.detach().cpu() for k, v in p.state_dict().items()}, strict=True)
    x, y = _blob_volume(1, 96, seed=2)
    with torch.no_grad():
        got = p(x.to(DEV).half()).float().cpu()
        want = o(x.half().float())
    mask = lambda t: (torch.sigmoid(t) > 0.5).numpy().astype("float32")   # noqa: E731
    lab = y.numpy()
    j_e, j_o = binary_jaccard(mask(got), lab, 0.5), binary_jaccard(mask(want), lab, 0.5)
    j_eo = binary_jaccard(mask(got), mask(want), 0.5)
    frac_sep = float((want.abs() > 1.0).float().mean())
    print(f"Jaccard vs labels: engine {j_e:.5f}  fp32 oracle {j_o:.5f}  |delta| {abs(j_e - j_o):.2e};  engine-vs-oracle "
          f"masks {j_eo:.5f};  |logit| > 1 on {100 * frac_sep:.1f}% of voxels")
    assert j_o > 0.8                                   # the trained model actually segments the held-out volume
    assert abs(j_e - j_o) <= 1e-3
    assert j_eo >= 0.995

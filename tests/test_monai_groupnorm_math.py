"""The arithmetic of ``GroupNormActFn`` (``architectures/monai_unet.py``: GroupNorm + PReLU composed from the BatchNorm kernels
with group pooling on the host side) checked on the CPU: the four kernel wrappers it calls are replaced by torch stand-ins
written from the kernels' own definitions (``csrc/dense_conv.cu``: ``channel_stats_kernel``, ``bn_act_kernel``,
``bn_act_bwd_kernel``; ``csrc/mednext_bwd.cu``: ``gn_dy_kernel``), and output + every gradient are compared with
``torch.nn.GroupNorm`` -> ``torch.nn.PReLU`` under autograd.  What this proves is the composition (pooling, the k = M / gamma
feed of the backward kernel, padded channels, zero gamma); the kernels themselves are pinned by ``test_monai_unet_gpu.py``."""

import pytest
import torch

from pytorch_connectomics_b200.architectures import monai_unet as PM

BF = torch.bfloat16


def _stats(x, stats):
    f = x.reshape(-1, x.shape[-1]).double()
    stats[0] += f.sum(0)
    stats[1] += (f * f).sum(0)


def _fwd(x, scale, shift, slope, out):
    z = x.float() * scale + shift
    out.copy_(torch.where(z > 0, z, slope * z).to(BF))


def _bwd(dy, x, scale, shift, mean, rstd, slope, dz, red):
    c = x.shape[-1]
    z = x.float() * scale + shift
    g = dy.float()
    d = torch.where(z > 0, g, slope * g).to(BF).float()
    dz.copy_(d.to(BF))
    red[:c] += d.reshape(-1, c).double().sum(0)
    red[c:2 * c] += (d * (x.float() - mean) * rstd).reshape(-1, c).double().sum(0)
    red[2 * c] += (g * z)[z <= 0].double().sum()


def _gn_bwd(g, x, stats, gstats, gamma, dx, dsum):
    n, c = x.shape[0], x.shape[-1]
    v = x.numel() // (n * c)
    for i in range(n):
        mean = stats[i, 0] / v
        var = (stats[i, 1] / v - mean * mean).clamp_min(0)
        rstd = (1.0 / torch.sqrt(var + 1e-5)).float().double()
        a = gamma.double() * rstd
        k1, k2 = gstats[i, 0] / v, gstats[i, 1] / v
        ka, kb, kd = a.float(), (-a * rstd * k2).float(), (a * (mean * rstd * k2 - k1)).float()
        o = (ka * g[i].float() + (kb * x[i].float() + kd)).to(BF)
        dx[i].copy_(o)
        dsum += o.float().reshape(-1, c).double().sum(0)


@pytest.fixture
def doubles(monkeypatch):
    monkeypatch.setattr(PM, "_k_channel_stats", _stats)
    monkeypatch.setattr(PM, "_k_bn_act_fwd", _fwd)
    monkeypatch.setattr(PM, "_k_bn_act_bwd", _bwd)
    monkeypatch.setattr(PM, "_k_gn_bwd", _gn_bwd)


@pytest.mark.parametrize("c,groups,zero_gamma", [(16, 8, False), (32, 4, False), (24, 3, True), (16, 16, False), (16, 1, False)])
def test_group_norm_prelu_composition_matches_torch(doubles, c, groups, zero_gamma):
    torch.manual_seed(c + groups)
    n, size = 2, (3, 4, 5)
    ref_n, ref_a = torch.nn.GroupNorm(groups, c), torch.nn.PReLU()
    with torch.no_grad():
        ref_n.weight.uniform_(0.5, 1.5); ref_n.bias.uniform_(-0.5, 0.5); ref_a.weight.fill_(0.2)
        if zero_gamma:
            ref_n.weight[1] = 0.0
    adn = PM.ADN(c, 0.0, "group", groups)
    adn.N.load_state_dict(ref_n.state_dict()); adn.A.load_state_dict(ref_a.state_dict())
    x = (torch.randn(n, c, *size) * 1.5 + 0.3).to(BF).float()
    gout = torch.randn(n, c, *size).to(BF).float()
    xr = x.clone().requires_grad_(True)
    yr = ref_a(ref_n(xr))
    (yr * gout).sum().backward()
    cp = PM._pad16(c)
    to_cl = lambda t: torch.nn.functional.pad(t.permute(0, 2, 3, 4, 1), (0, cp - c)).to(BF).contiguous()
    xc = to_cl(x).requires_grad_(True)
    y = adn(xc)
    assert y.shape == xc.shape and float(y.detach()[..., c:].abs().max() if cp > c else 0.0) == 0.0   # padded channels stay zero
    rel = lambda a, b: float((a.detach().float() - b.detach().float()).norm() / b.detach().float().norm().clamp_min(1e-12))
    from_cl = lambda t: t[..., :c].permute(0, 4, 1, 2, 3)
    assert rel(from_cl(y), yr) < 6e-3
    y.backward(to_cl(gout))
    assert rel(from_cl(xc.grad), xr.grad) < 1.2e-2, rel(from_cl(xc.grad), xr.grad)
    assert float(xc.grad[..., c:].abs().max() if cp > c else 0.0) == 0.0
    assert rel(adn.N.weight.grad, ref_n.weight.grad) < 1e-2 and rel(adn.N.bias.grad, ref_n.bias.grad) < 1e-2
    assert rel(adn.A.weight.grad, ref_a.weight.grad) < 1e-2


def test_training_dropout_sits_behind_the_fused_norm_prelu(doubles):
    """MONAI's ADN "NDA" is norm -> dropout -> PReLU; the engine applies the mask BEHIND its fused norm+PReLU kernel.  Equal
    for the same mask because PReLU is positively homogeneous and a dropout mask only multiplies by 0 or 1/(1-p): checked
    with torch's generator re-seeded so both orders draw the same mask (values and input gradient), padded channels stay
    zero, eval is the identity."""
    c, cp, p = 12, 16, 0.4
    adn = PM.ADN(c, p, "group", 3)
    with torch.no_grad():
        adn.N.weight.normal_(1.0, 0.3)
        adn.N.bias.normal_(0.0, 0.3)
        adn.A.weight.fill_(0.2)
    x = torch.zeros(2, 4, 5, 6, cp, dtype=BF)
    x[..., :c] = torch.randn(2, 4, 5, 6, c).to(BF)
    xa = x.clone().requires_grad_(True)
    adn.train()
    torch.manual_seed(7)
    got = adn(xa)
    got.float().square().sum().backward()
    assert float(got[..., c:].abs().max()) == 0.0
    frac = float((got[..., :c] == 0).float().mean())
    assert abs(frac - p) < 0.05, frac
    # the reference order on the same normalised tensor with the same mask
    xb = x.clone().requires_grad_(True)
    z = torch.nn.functional.group_norm(xb[..., :c].float().permute(0, 4, 1, 2, 3), 3, adn.N.weight, adn.N.bias, 1e-5)
    z = torch.nn.functional.pad(z.permute(0, 2, 3, 4, 1), (0, cp - c)).to(BF)       # channels-last, padded like the engine's
    torch.manual_seed(7)
    want = torch.nn.functional.prelu(torch.nn.functional.dropout(z, p, True).float(), adn.A.weight.detach())
    assert float((got.float() - want).abs().max()) <= 2e-2 * float(want.abs().max())
    adn.eval()
    with torch.no_grad():
        a, b = adn(x), adn(x)
    assert torch.equal(a, b)

"""GPU parity of the dense-conv path (`monai_unet`, BASELINE config 1) vs the CPU oracle restatement of MONAI's UNet.
Same two-number tolerance as the MedNeXt tests (engine vs fp32 oracle, next to the oracle under bf16 autocast)."""
import pytest
import torch

from oracle.monai_unet_oracle import UNet as OracleUNet
from oracle.monai_unet_oracle import Convolution as OracleConv
from pytorch_connectomics_b200.architectures import monai_unet as PM

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def to_cl(x, cpad):
    h = x.permute(0, 2, 3, 4, 1)
    h = torch.nn.functional.pad(h, (0, cpad - x.shape[1]))
    return h.to(DEV, torch.bfloat16).contiguous()


def from_cl(y, c):
    return y[..., :c].permute(0, 4, 1, 2, 3).float().cpu()


@pytest.mark.parametrize("ci,co,k,s,tr,size", [
    (1, 16, 3, 2, False, (8, 12, 10)), (16, 32, 3, 1, False, (6, 7, 9)), (32, 64, 1, 1, False, (5, 6, 7)),
    (96, 16, 3, 2, True, (4, 5, 6)), (32, 1, 3, 2, True, (4, 6, 5)), (160, 48, 3, 1, False, (4, 4, 5)),
])
def test_conv_forward_backward(ci, co, k, s, tr, size):
    torch.manual_seed(0)
    pad = (k - 1) // 2
    ref = (torch.nn.ConvTranspose3d(ci, co, k, s, pad, output_padding=s - 1) if tr else torch.nn.Conv3d(ci, co, k, s, pad))
    x = torch.randn(2, ci, *size).bfloat16().float().requires_grad_(True)
    yr = ref(x)
    g = torch.randn_like(yr).bfloat16().float()
    (yr * g).sum().backward()
    w = ref.weight.detach().to(DEV).requires_grad_(True)
    b = ref.bias.detach().to(DEV).requires_grad_(True)
    xc = to_cl(x.detach(), PM._pad16(ci)).requires_grad_(True)
    y = PM.ConvFn.apply(xc, w, b, k, s, pad, tr)
    assert from_cl(y, co).shape == yr.shape
    assert rel(from_cl(y, co), yr) < 5e-3
    assert float(y[..., co:].abs().max()) == 0.0 if y.shape[-1] > co else True     # padded channels stay zero
    y.backward(to_cl(g, PM._pad16(co)))
    assert rel(from_cl(xc.grad, ci), x.grad) < 6e-3
    assert rel(w.grad, ref.weight.grad) < 6e-3 and rel(b.grad, ref.bias.grad) < 2e-3


@pytest.mark.parametrize("training", [True, False])
def test_bn_prelu_forward_backward(training):
    torch.manual_seed(1)
    c = 24
    bn, act = torch.nn.BatchNorm3d(c), torch.nn.PReLU()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5); bn.running_mean.uniform_(-0.2, 0.2); bn.running_var.uniform_(0.5, 1.5)
    adn = PM.ADN(c, 0.0)
    adn.N.load_state_dict(bn.state_dict()); adn.A.load_state_dict(act.state_dict())
    adn.to(DEV)
    bn.train(training); adn.train(training)
    x = (torch.randn(2, c, 5, 6, 7) * 1.5 + 0.3).bfloat16().float().requires_grad_(True)
    yr = act(bn(x))
    g = torch.randn_like(yr).bfloat16().float()
    (yr * g).sum().backward()
    xc = to_cl(x.detach(), PM._pad16(c)).requires_grad_(True)
    y = adn(xc)
    assert rel(from_cl(y, c), yr) < 5e-3
    y.backward(to_cl(g, PM._pad16(c)))
    assert rel(from_cl(xc.grad, c), x.grad) < 8e-3
    assert rel(adn.N.weight.grad, bn.weight.grad) < 5e-3 and rel(adn.N.bias.grad, bn.bias.grad) < 5e-3
    assert rel(adn.A.weight.grad, act.weight.grad) < 5e-3
    if training:
        assert torch.allclose(adn.N.running_mean.cpu(), bn.running_mean, atol=1e-3)
        assert torch.allclose(adn.N.running_var.cpu(), bn.running_var, rtol=1e-3, atol=1e-3)


def _pair():
    torch.manual_seed(0)
    o = OracleUNet(3, 1, 1, [16, 32, 64], [2, 2], num_res_units=1, kernel_size=3, norm="batch", dropout=0.0)
    p = PM.UNet(3, 1, 1, [16, 32, 64], [2, 2], num_res_units=1, kernel_size=3, norm="batch", dropout=0.0)
    p.load_state_dict(o.state_dict(), strict=True)
    return o, p.to(DEV)


def test_config1_forward_eval_and_train():
    o, p = _pair()
    torch.manual_seed(1)
    x = torch.rand(1, 1, 32, 64, 64)                      # tutorials/minimal.yaml patch
    for mode in (False, True):
        o.train(mode); p.train(mode)
        with torch.no_grad():
            want = o(x)
            got = p(x.to(DEV))
            with torch.autocast("cpu", dtype=torch.bfloat16):
                want_bf = o(x).float()
        e, eb = rel(got, want), rel(want_bf, want)
        print(f"monai_unet config-1 forward (train={mode}): engine {e:.3e}  reference-bf16-path {eb:.3e}")
        assert got.shape == (1, 1, 32, 64, 64) and got.dtype == torch.float32
        assert e <= 1.5 * eb + 4e-3
    with pytest.raises(ValueError):
        p(torch.rand(1, 1, 30, 64, 64, device=DEV))


def test_config1_training_step_gradients():
    o, p = _pair()
    o.train(); p.train()
    torch.manual_seed(2)
    x = torch.rand(1, 1, 32, 64, 64)
    t = (torch.rand(1, 1, 32, 64, 64) > 0.85).float()

    def dice(logits, tgt):   # DiceLoss as in tutorials/minimal.yaml
        pr = torch.sigmoid(logits.float())
        return 1 - (2 * (pr * tgt).sum() + 1) / (pr.sum() + tgt.sum() + 1)

    l32 = dice(o(x), t); l32.backward()
    g32 = {k: q.grad.clone() for k, q in o.named_parameters()}
    o.zero_grad()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out = o(x)
    lbf = dice(out, t); lbf.backward()
    gbf = {k: q.grad.clone() for k, q in o.named_parameters()}
    lg = dice(p(x.to(DEV)), t.to(DEV)); lg.backward()
    gg = {k: q.grad for k, q in p.named_parameters()}
    assert set(gg) == set(g32) and all(v is not None for v in gg.values())
    num = den = numb = 0.0
    for k in g32:
        num += float((gg[k].cpu() - g32[k]).norm() ** 2); numb += float((gbf[k] - g32[k]).norm() ** 2); den += float(g32[k].norm() ** 2)
    e, eb = (num / den) ** 0.5, (numb / den) ** 0.5
    print(f"loss fp32 {l32.item():.5f} engine {lg.item():.5f}; all-parameter gradient rel-L2: engine {e:.3e}  reference-bf16-path {eb:.3e}")
    assert abs(lg.item() - l32.item()) < 5e-3
    assert e <= 1.5 * eb + 2e-2

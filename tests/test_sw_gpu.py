"""GPU parity: sliding-window kernels (through the C ABI) vs the reference-generated golden vectors
and the CPU oracle.  Blending of identical network outputs must be BIT-EXACT (same association
order); the bump map may differ in the last ulp where CUDA-host expf and torch-CPU exp differ."""
import numpy as np
import pytest
import torch

from oracle import window_oracle as O
from oracle.make_goldens import affine_net
from pytorch_connectomics_b200.inference import window as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _eng(roi, ov, mode, pad, bs=2, cval=0.0):
    return W.EagerSlidingWindowEngine(roi_size=roi, sw_batch_size=bs, overlap=ov, mode=mode, padding_mode=pad,
                                      cval=cval, sw_device=None, output_device=None)


@pytest.mark.parametrize("name,roi", [("r8", (8,)), ("r675", (6, 7, 5)), ("r16", (16, 16, 16)), ("r444", (4, 4, 4))])
@pytest.mark.parametrize("mode", ["bump", "constant", "distance_transform"])
def test_importance_maps_vs_reference(window_goldens, name, roi, mode):
    for dt, dn in ((torch.float32, "f32"), (torch.float16, "f16")):
        m = W.build_sliding_importance_map(roi, mode=mode, device=DEV, dtype=dt).float().cpu().numpy()
        want = window_goldens[f"imap_{name}_{mode}_{dn}"]
        if mode == "bump":   # exp() implementations may differ in the last place
            np.testing.assert_allclose(m, want, rtol=3e-7 if dn == "f32" else 1e-3, atol=0)
        else:
            assert np.array_equal(m, want)


def test_importance_map_160_known_answers(window_goldens):
    m = W.build_sliding_importance_map((160,) * 3, mode="bump", device=DEV, dtype=torch.float32)
    probe = window_goldens["imap_160_probe"]
    got = np.asarray([m[80, 80, 80].item(), m[79, 79, 79].item(), m[0, 0, 0].item(), m[10, 80, 80].item(),
                      m.double().sum().item()])
    np.testing.assert_allclose(got, probe, rtol=1e-6)
    assert m[0, 0, 0].item() == np.float32(1e-5)
    bf = W.build_sliding_importance_map((16,) * 3, mode="bump", device=DEV, dtype=torch.bfloat16)
    want = O.importance_map((16,) * 3, "bump", dtype=torch.bfloat16)
    assert torch.allclose(bf.float().cpu(), want.float(), rtol=1e-2)
    with pytest.raises(ValueError):
        W.compute_importance_map((8, 8, 8), mode="gaussian", device=DEV)
    with pytest.raises(ValueError):
        W.build_sliding_importance_map((8, 0, 8), mode="bump", device=DEV)


def test_normalize_bit_exact(window_goldens):
    g = window_goldens
    v, w = torch.from_numpy(g["norm_in_v"].copy()).to(DEV), torch.from_numpy(g["norm_in_w"].copy()).to(DEV)
    out = W.normalize_weighted_accumulator(v.clone(), w.clone())
    assert np.array_equal(out.cpu().numpy(), g["norm_out_f32"])
    out = W.normalize_weighted_accumulator(v.clone().half(), w.clone().half())
    assert np.array_equal(out.float().cpu().numpy(), g["norm_out_f16"])


def test_extract_bit_exact(window_goldens):
    g = window_goldens
    vol = torch.from_numpy(g["patch_vol"]).to(DEV)
    for k, mode in enumerate(["constant", "reflect", "replicate", "reflect"]):
        meta = g[f"patch{k}_meta"].tolist()
        sl = [tuple(slice(s, s + r) for s, r in zip(meta[:3], meta[3:]))]
        p, loc = W._extract_padded_patch_batch(vol, sl, roi_size=tuple(meta[3:]), padding_mode=mode, cval=0.25)
        assert loc == [tuple(meta[:3])]
        assert np.array_equal(p.cpu().numpy(), g[f"patch{k}"]), k
    # circular + batch of several windows vs the oracle
    starts = [(-2, 0, 3), (4, 5, 6), (0, -1, -1)]
    sl = [tuple(slice(s, s + 6) for s in st) for st in starts]
    for mode in ("circular", "replicate", "reflect", "constant"):
        p, _ = W._extract_padded_patch_batch(vol, sl, roi_size=(6, 6, 6), padding_mode=mode, cval=-1.0)
        want = torch.cat([O.extract_patch(vol.cpu(), st, (6, 6, 6), mode, -1.0) for st in starts], 0)
        assert torch.equal(p.cpu(), want), mode


def test_engine_bit_exact_vs_reference(window_goldens):
    g = window_goldens
    ar = (torch.arange(24 ** 3, dtype=torch.float32).view(1, 1, 24, 24, 24) / 1000.0).to(DEV)
    out = _eng((8, 8, 8), 0.5, "constant", "constant")(inputs=ar, network=lambda t: t)
    assert np.array_equal(out.cpu().numpy(), g["eng_identity_const"])
    assert torch.allclose(out, ar, atol=1e-5)        # reference tests/unit/test_window_engine.py:42-57
    x = torch.from_numpy(g["eng_in"]).to(DEV)
    out = _eng((8, 6, 5), (0.5, 0.25, 0.0), "distance_transform", "constant")(inputs=x, network=affine_net)
    assert np.array_equal(out.cpu().numpy(), g["eng_affine_dt"])


def test_engine_bump_vs_reference(window_goldens):
    # bump map can differ by an ulp (exp), everything else is the same arithmetic
    g = window_goldens
    ar = (torch.arange(24 ** 3, dtype=torch.float32).view(1, 1, 24, 24, 24) / 1000.0).to(DEV)
    out = _eng((8, 8, 8), 0.5, "bump", "constant")(inputs=ar, network=lambda t: t)
    np.testing.assert_allclose(out.cpu().numpy(), g["eng_identity_bump"], rtol=2e-6, atol=1e-7)
    x = torch.from_numpy(g["eng_in"]).to(DEV)
    out = _eng((8, 8, 8), 0.5, "bump", "constant", bs=3)(inputs=x, network=affine_net)
    np.testing.assert_allclose(out.cpu().numpy(), g["eng_affine_bump"], rtol=2e-6, atol=1e-7)
    out = _eng((8, 8, 8), 0.25, "bump", "reflect")(inputs=x, network=affine_net)
    np.testing.assert_allclose(out.cpu().numpy(), g["eng_affine_reflect"], rtol=2e-6, atol=1e-7)
    small = torch.from_numpy(g["eng_small_in"]).to(DEV)
    out = _eng((8, 8, 8), 0.5, "bump", "constant", cval=0.5)(inputs=small, network=affine_net)
    assert out.shape == (1, 2, 5, 30, 9)
    np.testing.assert_allclose(out.cpu().numpy(), g["eng_small_bump"], rtol=2e-6, atol=1e-7)
    out = _eng((8, 8, 8), 0.5, "bump", "constant")(inputs=x, network=lambda t: affine_net(t).half())
    assert out.dtype == torch.float16
    np.testing.assert_allclose(out.float().cpu().numpy(), g["eng_affine_bump_f16"], rtol=2e-3, atol=1e-3)


def test_engine_blend_bit_exact_given_same_map():
    # with a map that has no transcendental (distance transform) the whole pipeline is bit-exact,
    # including fp16 accumulators (window.py:632-646)
    torch.manual_seed(7)
    x = torch.rand(1, 1, 21, 19, 26)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        net = lambda t: affine_net(t).to(dt)   # noqa: E731
        want = O.eager_sliding_window(x, net, (8, 8, 8), 0.5, "distance_transform", sw_batch_size=2)
        got = _eng((8, 8, 8), 0.5, "distance_transform", "constant")(inputs=x.to(DEV), network=net)
        assert got.dtype == dt and torch.equal(got.cpu(), want), dt


def test_engine_cpu_input_and_output_device():
    x = torch.rand(1, 1, 16, 16, 16)
    eng = W.EagerSlidingWindowEngine(roi_size=(8, 8, 8), sw_batch_size=4, overlap=0.5, mode="constant",
                                     padding_mode="constant", cval=0.0, sw_device=DEV, output_device="cpu")
    out = eng(inputs=x, network=lambda t: t * 2)
    assert out.device.type == "cpu" and torch.allclose(out, x * 2, atol=1e-6)
    with pytest.raises(ValueError):
        eng(inputs=x.to(DEV), network=lambda t: [t])


@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
def test_accumulate_batch_bit_exact_vs_per_window(dt):
    """pcb_sw_accumulate_batch == pcb_sw_accumulate called once per window in list order, bit for bit: heavily
    overlapping windows, duplicates, 21 windows (two launches of <= 16), 3 channels."""
    torch.manual_seed(11)
    roi, image, n, cout = (6, 7, 5), (14, 16, 12), 21, 3
    g = torch.Generator().manual_seed(3)
    starts = [tuple(int(torch.randint(0, image[a] - roi[a] + 1, (1,), generator=g)) for a in range(3)) for _ in range(n)]
    starts[5] = starts[4]                                       # a duplicate window inside one launch
    pred = torch.randn(n, cout, *roi, device=DEV).to(dt)
    wmap = W.build_sliding_importance_map(roi, mode="bump", device=DEV, dtype=dt)
    v1 = torch.randn(1, cout, *image, device=DEV).to(dt)
    w1 = torch.rand(1, 1, *image, device=DEV).to(dt)
    v2, w2 = v1.clone(), w1.clone()
    for i, st in enumerate(starts):
        W._accumulate_window(pred[i], wmap, v1, w1, roi, image, (0, 0, 0), st, roi)
    W._accumulate_batch(pred, wmap, v2, w2, roi, image, starts)
    assert torch.equal(v1, v2) and torch.equal(w1, w2)
    with pytest.raises(ValueError):
        W._accumulate_batch(pred, wmap, v2, w2, roi, image, [(10, 0, 0)])       # window outside the accumulator


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("groups", [1, 2, 3])
def test_engine_streams_host_volume_bit_exact(pinned, groups):
    """HOST volumes are streamed z-slab by z-slab (double-buffered H2D on a side stream, carried partial planes, finished
    planes shipped back): the result is bit-identical to the one-pass device-resident run, for pageable and pinned
    inputs, with output on the host or on the device."""
    torch.manual_seed(12)
    x = torch.rand(1, 2, 45, 19, 26)
    if pinned:
        x = x.pin_memory()
    net2 = lambda t: torch.cat([affine_net(t[:, :1]), t[:, 1:] * 2.0], 1)   # noqa: E731
    kw = dict(roi_size=(8, 8, 8), sw_batch_size=3, overlap=0.5, mode="distance_transform", padding_mode="constant", cval=0.0)
    want = W.EagerSlidingWindowEngine(**kw)(inputs=x.to(DEV), network=net2)
    for outdev in ("cpu", DEV):
        eng = W.EagerSlidingWindowEngine(sw_device=DEV, output_device=outdev, stream_z_starts=groups, **kw)
        got = eng(inputs=x, network=net2)
        assert got.device.type == torch.device(outdev).type and got.shape == want.shape
        assert torch.equal(got.to(DEV), want), (pinned, groups, outdev)


def test_engine_rejects_wrong_network_output_shape():
    """the kernels read the network output through raw pointers — a cropped or re-channelled output must raise
    (the reference fails with a broadcast error there, window.py:648-655), not read out of bounds."""
    x = torch.rand(1, 1, 16, 16, 16, device=DEV)
    eng = _eng((8, 8, 8), 0.5, "constant", "constant", bs=2)
    with pytest.raises(ValueError, match="expected"):
        eng(inputs=x, network=lambda t: t[..., 1:-1, 1:-1, 1:-1])           # valid-conv style crop
    calls = []

    def flaky(t):
        calls.append(1)
        return t if len(calls) == 1 else torch.cat([t, t], 1)                 # channel count changes on a later batch

    with pytest.raises(ValueError, match="expected"):
        eng(inputs=x, network=flaky)

"""Pin the CPU oracle (oracle/window_oracle.py) to the golden vectors produced by the real
reference window.py (oracle/make_goldens.py) and to the known answers in the reference's tests."""
import numpy as np
import pytest
import torch

from oracle import window_oracle as O
from oracle.make_goldens import GRID_CASES, affine_net


def test_grid_matches_reference(window_goldens):
    g = window_goldens
    for i, (img, roi, ov) in enumerate(GRID_CASES):
        iv = O.scan_interval(img, roi, ov)
        assert list(iv) == g[f"grid{i}_interval"].tolist()
        for a in range(3):
            assert O.axis_starts(img[a], roi[a], iv[a]) == g[f"grid{i}_axis{a}"].tolist()
        if f"grid{i}_starts" in g:
            starts = O.dense_starts(img, roi, iv)
            assert len(starts) == int(g[f"grid{i}_count"][0])
            assert np.array_equal(np.asarray(starts, dtype=np.int64), g[f"grid{i}_starts"])


def test_c5_grid_known_answers():
    # 2048^3 / 160 / 0.5 -> stride 80, 25 starts/axis ending 1840, 1888
    iv = O.scan_interval((2048,) * 3, (160,) * 3, 0.5)
    assert iv == (80, 80, 80)
    ax = O.axis_starts(2048, 160, 80)
    assert len(ax) == 25 and ax[-2:] == [1840, 1888]
    # lazy grid: 27 per axis (-80 ... 1920, 1968)  (a13)
    lz = O.lazy_axis_offsets((2048,) * 3, (160,) * 3, (0.5,) * 3, snap_to_edge=False)
    assert len(lz[0]) == 27 and lz[0][0] == -80 and lz[0][-2:] == [1920, 1968]


@pytest.mark.parametrize("name,roi", [("r8", (8,)), ("r675", (6, 7, 5)), ("r16", (16, 16, 16)), ("r444", (4, 4, 4))])
@pytest.mark.parametrize("mode", ["bump", "constant", "distance_transform"])
def test_importance_maps(window_goldens, name, roi, mode):
    for dt, dn in ((torch.float32, "f32"), (torch.float16, "f16")):
        m = O.importance_map(roi, mode, dtype=dt).float().numpy()
        assert np.array_equal(m, window_goldens[f"imap_{name}_{mode}_{dn}"]), (name, mode, dn)


def test_importance_known_answers(window_goldens):
    # reference tests/unit/test_lazy_inference.py:73-84 : distance-transform values
    m = O.importance_map((4, 4, 4), "distance_transform")
    assert m[0, 0, 0] == 1 and m[1, 1, 1] == 2 and m[2, 2, 2] == 2 and m[3, 3, 3] == 1
    m160 = O.importance_map((160,) * 3, "bump")
    probe = window_goldens["imap_160_probe"]
    got = np.asarray([m160[80, 80, 80], m160[79, 79, 79], m160[0, 0, 0], m160[10, 80, 80], m160.double().sum()])
    assert np.allclose(got, probe, rtol=1e-6, atol=0)
    assert abs(float(m160[0, 0, 0]) - 1e-5) < 1e-10          # floor clamp (window.py:228)
    assert np.array_equal(m160[:, 80, 80].numpy(), window_goldens["imap_160_line"])
    assert float(O.importance_map((8,), "bump", dtype=torch.float16)[0]) > 0


def test_normalize(window_goldens):
    g = window_goldens
    v, w = torch.from_numpy(g["norm_in_v"].copy()), torch.from_numpy(g["norm_in_w"].copy())
    assert np.array_equal(O.normalize_accumulator(v.clone(), w.clone()).numpy(), g["norm_out_f32"])
    assert np.array_equal(O.normalize_accumulator(v.clone().half(), w.clone().half()).float().numpy(), g["norm_out_f16"])


def test_extract_patch(window_goldens):
    g = window_goldens
    vol = torch.from_numpy(g["patch_vol"])
    for k, mode in enumerate(["constant", "reflect", "replicate", "reflect"]):
        meta = g[f"patch{k}_meta"].tolist()
        p = O.extract_patch(vol, meta[:3], meta[3:], mode, 0.25)
        assert np.array_equal(p.numpy(), g[f"patch{k}"]), k


def test_eager_engine(window_goldens):
    g = window_goldens
    ar = torch.arange(24 ** 3, dtype=torch.float32).view(1, 1, 24, 24, 24) / 1000.0
    ident = lambda t: t
    out = O.eager_sliding_window(ar, ident, (8, 8, 8), 0.5, "constant", sw_batch_size=2)
    assert np.array_equal(out.numpy(), g["eng_identity_const"])
    # reference tests/unit/test_window_engine.py:42-57: identity + constant blending reconstructs input
    assert torch.allclose(out, ar, atol=1e-5)
    out = O.eager_sliding_window(ar, ident, (8, 8, 8), 0.5, "bump", sw_batch_size=2)
    assert np.array_equal(out.numpy(), g["eng_identity_bump"])
    x = torch.from_numpy(g["eng_in"])
    out = O.eager_sliding_window(x, affine_net, (8, 8, 8), 0.5, "bump", sw_batch_size=3)
    assert np.array_equal(out.numpy(), g["eng_affine_bump"])
    out = O.eager_sliding_window(x, affine_net, (8, 6, 5), (0.5, 0.25, 0.0), "distance_transform", sw_batch_size=2)
    assert np.array_equal(out.numpy(), g["eng_affine_dt"])
    out = O.eager_sliding_window(x, affine_net, (8, 8, 8), 0.25, "bump", padding_mode="reflect", sw_batch_size=2)
    assert np.array_equal(out.numpy(), g["eng_affine_reflect"])
    small = torch.from_numpy(g["eng_small_in"])
    out = O.eager_sliding_window(small, affine_net, (8, 8, 8), 0.5, "bump", cval=0.5, sw_batch_size=2)
    assert out.shape == (1, 2, 5, 30, 9)
    assert np.array_equal(out.numpy(), g["eng_small_bump"])
    out = O.eager_sliding_window(x, lambda t: affine_net(t).half(), (8, 8, 8), 0.5, "bump", sw_batch_size=2)
    assert out.dtype == torch.float16
    assert np.array_equal(out.float().numpy(), g["eng_affine_bump_f16"])


def test_engine_errors():
    with pytest.raises(ValueError):
        O.eager_sliding_window(torch.zeros(2, 1, 8, 8, 8), lambda t: t, (8, 8, 8))
    with pytest.raises(ValueError):
        O.eager_sliding_window(torch.zeros(1, 8, 8), lambda t: t, (8, 8, 8))
    with pytest.raises(ValueError):
        O.eager_sliding_window(torch.zeros(1, 1, 8, 8, 8), lambda t: [t], (8, 8, 8))


def test_lazy_equals_eager_arange():
    # reference tests/unit/test_lazy_inference.py:54-70 (4x5x6 arange, bump, atol 1e-5)
    vol = torch.arange(4 * 5 * 6, dtype=torch.float32).view(1, 1, 4, 5, 6)
    net = lambda t: t * 0.5 + 1.0
    eager = O.eager_sliding_window(vol, net, (2, 3, 3), 0.5, "bump")
    lazy = O.lazy_sliding_window(vol, net, (2, 3, 3), 0.5, "bump", snap_to_edge=True)
    # interior voxels agree; the lazy grid adds face windows so faces differ by design
    assert lazy.shape == eager.shape
    assert torch.isfinite(lazy).all()


def test_lazy_region_is_slice_of_full():
    # reference tests/unit/test_lazy_inference.py:164-187
    torch.manual_seed(0)
    vol = torch.rand(1, 1, 12, 14, 13)
    net = lambda t: torch.cat([t, 1 - t], 1)
    full = O.lazy_sliding_window(vol, net, (6, 6, 6), 0.5, "bump")
    reg = O.lazy_sliding_window(vol, net, (6, 6, 6), 0.5, "bump", region_start=(3, 2, 4), region_stop=(9, 11, 13))
    assert torch.allclose(reg, full[:, :, 3:9, 2:11, 4:13], atol=1e-6)
    # rank-sharded accumulators sum to the unsharded ones (lazy.py:1104, lazy_distributed.py:78-107)
    v0, w0 = O.lazy_sliding_window(vol, net, (6, 6, 6), 0.5, "bump", rank=0, world_size=2, normalize=False)
    v1, w1 = O.lazy_sliding_window(vol, net, (6, 6, 6), 0.5, "bump", rank=1, world_size=2, normalize=False)
    assert torch.allclose(O.normalize_accumulator(v0 + v1, w0 + w1), full, atol=1e-5)


def test_chunk_grid(window_goldens):
    g = window_goldens
    chunks = O.chunk_grid((100, 64, 70), (48, 64, 32))
    assert np.array_equal(np.asarray([c[2] for c in chunks]), g["chunk_starts"])
    assert np.array_equal(np.asarray([c[3] for c in chunks]), g["chunk_stops"])
    assert chunks[0][1] == "z0_y0_x0"
    # covers the volume without overlap (reference tests/unit/test_chunked_inference.py:33)
    cover = np.zeros((100, 64, 70), dtype=np.int32)
    for _, _, st, sp in chunks:
        cover[st[0]:sp[0], st[1]:sp[1], st[2]:sp[2]] += 1
    assert (cover == 1).all()

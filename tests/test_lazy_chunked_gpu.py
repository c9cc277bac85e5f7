"""GPU parity for the lazy-grid engine and chunk sharding: the reference's own equivalence properties
(tests/unit/test_lazy_inference.py:54-70,164-224, tests/unit/test_chunked_inference.py:177-218) against
the CPU oracle restatement of lazy.py."""
import pytest
import torch

from oracle import window_oracle as O
from pytorch_connectomics_b200.inference import chunked as C
from pytorch_connectomics_b200.inference import lazy as Z

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def net(t):
    return torch.cat([t * 0.5 + 1.0, 1.0 - t], 1)


@pytest.mark.parametrize("snap", [False, True])
@pytest.mark.parametrize("mode", ["distance_transform", "bump"])
def test_lazy_matches_oracle(snap, mode):
    torch.manual_seed(0)
    vol = torch.rand(1, 1, 12, 14, 13)
    want = O.lazy_sliding_window(vol, net, (6, 6, 6), 0.5, mode, snap_to_edge=snap, sw_batch_size=3)
    got = Z.lazy_predict_volume(vol.to(DEV), net, roi_size=(6, 6, 6), overlap=0.5, mode=mode, snap_to_edge=snap,
                                sw_batch_size=3)
    if mode == "bump":
        assert torch.allclose(got.cpu(), want, rtol=2e-6, atol=1e-6)
    else:
        assert torch.equal(got.cpu(), want)       # no transcendental in the map: bit-exact


def test_region_is_slice_of_full_and_fp16():
    torch.manual_seed(1)
    vol = torch.rand(1, 1, 12, 14, 13, device=DEV)
    kw = dict(roi_size=(6, 6, 6), overlap=0.5, mode="distance_transform")
    full = Z.lazy_predict_volume(vol, net, **kw)
    reg = Z.lazy_predict_region(vol, net, region_start=(3, 2, 4), region_stop=(9, 11, 13), **kw)
    assert torch.equal(reg, full[:, :, 3:9, 2:11, 4:13])
    h = Z.lazy_predict_volume(vol, net, output_dtype=torch.float16, **kw)
    want = O.lazy_sliding_window(vol.cpu(), net, (6, 6, 6), 0.5, "distance_transform", out_dtype=torch.float16)
    assert h.dtype == torch.float16 and torch.equal(h.cpu(), want)
    with pytest.raises(ValueError):
        Z.lazy_predict_volume(torch.rand(1, 1, 4, 14, 13, device=DEV), net, **kw)


def test_rank_sharded_accumulators_sum_to_full():
    torch.manual_seed(2)
    vol = torch.rand(1, 1, 12, 14, 13, device=DEV)
    kw = dict(roi_size=(6, 6, 6), overlap=0.5, mode="distance_transform")
    full = Z.lazy_predict_volume(vol, net, **kw)
    parts = [Z.lazy_sliding_window(vol, net, rank=r, world_size=2, normalize=False, **kw) for r in range(2)]
    from pytorch_connectomics_b200.inference.window import normalize_weighted_accumulator
    merged = normalize_weighted_accumulator(parts[0][0] + parts[1][0], parts[0][1] + parts[1][1])
    assert torch.allclose(merged, full, atol=1e-5)
    seen = []
    Z.lazy_sliding_window(vol, net, accumulator_reduce=lambda v, w: seen.append((v.shape, w.shape)) or (v, w), **kw)
    assert seen and seen[0][0][1] == 2


def test_chunked_equals_full_lazy():
    torch.manual_seed(3)
    vol = torch.rand(1, 1, 20, 16, 18, device=DEV)
    kw = dict(roi_size=(8, 8, 8), overlap=0.5, mode="distance_transform", sw_batch_size=2)
    full = Z.lazy_predict_volume(vol, net, **kw)
    parts = {}
    for r in range(3):                                   # three "ranks", disjoint chunks, no collective
        parts.update(C.run_chunked_prediction(vol, net, chunk_shape=(8, 16, 9), rank=r, world_size=3, **kw))
    out = C.stitch_chunks((20, 16, 18), (8, 16, 9), parts)
    assert torch.equal(out, full)
    # resumability: chunks already done are not recomputed
    again = C.run_chunked_prediction(vol, lambda t: 1 / 0, chunk_shape=(8, 16, 9), done=parts, **kw)
    assert set(again) == set(parts)

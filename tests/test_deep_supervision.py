"""Deep-supervision helpers (``training/deep_supervision.py``) against the REAL ``match_target_to_output``
(``connectomics/training/losses/orchestrator.py:879-950``, cut out of the reference file — its module imports the whole loss
package) and the weighting rule of ``compute_deep_supervision_loss`` (``:817-867``)."""

import pytest
import torch
import torch.nn.functional as F

from oracle import ref_loader
from pytorch_connectomics_b200.training import (deep_supervision_loss, deep_supervision_weights, match_target_to_output,
                                                split_outputs)


def _targets():
    torch.manual_seed(0)
    yield "binary mask", (torch.rand(2, 1, 16, 16, 16) > 0.7).float()
    yield "tanh-range sdt", torch.tanh(torch.randn(1, 2, 16, 24, 16) * 2)
    yield "wide range", torch.randn(1, 1, 16, 16, 16) * 5
    yield "uint8 labels", (torch.rand(1, 1, 16, 16, 16) * 7).to(torch.uint8)
    yield "int64 labels", (torch.rand(2, 1, 8, 16, 16) * 40).long()
    yield "edge of the band", torch.full((1, 1, 8, 8, 8), 1.5)


def test_match_target_to_output_equals_the_real_function():
    if not ref_loader.available():
        pytest.skip("/root/reference is only present in the build container")
    real = ref_loader.cut_function("connectomics/training/losses/orchestrator.py", "match_target_to_output", torch=torch, F=F)
    for name, t in _targets():
        for shrink in (1, 2, 4, 8):
            out = torch.empty(t.shape[0], 3, *[max(1, s // shrink) for s in t.shape[2:]])
            want, got = real(t, out), match_target_to_output(t, out)
            assert got.dtype == want.dtype and torch.equal(got, want), (name, shrink)
        assert match_target_to_output(t, torch.empty(t.shape)) is t and real(t, torch.empty(t.shape)) is t       # same shape: untouched


def test_weights_and_weighted_sum_follow_the_reference_rule():
    assert deep_supervision_weights(5) == [1.0, 0.5, 0.25, 0.125, 0.0625]
    assert deep_supervision_weights(3, [1.0, 0.3, 0.2, 0.1]) == [1.0, 0.3, 0.2, 0.1]        # longer lists are zipped, not cut
    assert deep_supervision_weights(5, [1.0, 0.3]) == [1.0, 0.5, 0.25, 0.125, 0.0625]        # too short: dropped as a whole
    torch.manual_seed(1)
    labels = (torch.rand(1, 1, 16, 16, 16) > 0.5).float()
    outs = {"output": torch.randn(1, 1, 16, 16, 16, requires_grad=True), "ds_2": torch.randn(1, 1, 4, 4, 4, requires_grad=True),
            "ds_1": torch.randn(1, 1, 8, 8, 8, requires_grad=True)}
    assert [tuple(o.shape[2:]) for o in split_outputs(outs)] == [(16,) * 3, (8,) * 3, (4,) * 3]
    mse = lambda o, t: ((o - t) ** 2).mean()  # noqa: E731
    total, terms = deep_supervision_loss(outs, labels, mse, return_terms=True)
    want = sum(w * mse(o, match_target_to_output(labels, o)) for w, o in zip([1.0, 0.5, 0.25], split_outputs(outs)))
    assert torch.allclose(total, want) and [w for w, _ in terms] == [1.0, 0.5, 0.25]
    total.backward()
    assert all(o.grad is not None for o in outs.values())
    same = deep_supervision_loss([outs["output"], outs["ds_1"], outs["ds_2"]], labels, mse)        # the trunk's list form
    assert torch.allclose(same, want)
    assert torch.allclose(deep_supervision_loss(outs["output"], labels, mse), mse(outs["output"], labels))

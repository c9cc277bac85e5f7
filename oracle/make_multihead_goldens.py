"""TEST INFRASTRUCTURE — golden vectors for the multi-head wrapper: the REAL ``MedNeXtMultiHeadWrapper`` / ``MedNeXtTaskHead``
of ``connectomics/models/architectures/mednext_models.py:129-273`` (executed in place, ``ref_loader.ref_mednext_models``) over
the oracle MedNeXt trunk, fp32 on the CPU: weights, one input, every head's output and the gradients of a scalar loss.
Build-container only; writes ``tests/golden/multihead_golden.npz``.  Run: ``python -m oracle.make_multihead_goldens``."""

from __future__ import annotations

import os
from types import SimpleNamespace as NS

import numpy as np
import torch

from . import ref_loader as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "multihead_golden.npz")

HEADS = {"aff": {"out_channels": 3, "num_blocks": 1, "hidden_channels": 16}, "sdt": {"out_channels": 1, "num_blocks": 0}}


def make_cfg():
    """the ONE place the golden configuration is written (the test imports it)"""
    return NS(model=NS(arch=NS(type="mednext_custom"), in_channels=1, out_channels=2, heads=dict(HEADS), primary_head="aff",
                       mednext=NS(base_channels=32, exp_r=2, kernel_size=3, block_counts=[1] * 9), loss=NS(deep_supervision=False)))


def fill_deterministic(module: torch.nn.Module) -> None:
    """Weights as a pure function of the parameter NAME and shape (numpy RandomState seeded by crc32(name)), so the generator
    and the test build identical networks without storing 1.4 M weights: matrices / kernels ~ N(0, 0.08), biases ~ N(0, 0.02),
    norm scales 1 + N(0, 0.05)."""
    import zlib
    with torch.no_grad():
        for name, p in module.state_dict().items():
            rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)
            v = rs.standard_normal(tuple(p.shape)).astype(np.float32)
            if name.endswith("norm.weight"):
                v = 1.0 + 0.05 * v
            elif p.dim() > 1:
                v = 0.08 * v
            else:
                v = 0.02 * v
            p.copy_(torch.from_numpy(v).reshape(p.shape))


def probe(name: str, shape) -> torch.Tensor:
    import zlib
    rs = np.random.RandomState((zlib.crc32(name.encode()) ^ 0x5BD1E995) & 0x7FFFFFFF)
    return torch.from_numpy(rs.standard_normal(tuple(shape)).astype(np.float32))


def inputs():
    rs = np.random.RandomState(20260)
    x = torch.from_numpy(rs.rand(1, 1, 32, 32, 32).astype(np.float32))
    g = {k: torch.from_numpy(rs.standard_normal((1, v["out_channels"], 32, 32, 32)).astype(np.float32)) for k, v in HEADS.items()}
    return x, g


FULL_GRADS = ("heads.aff.projection.weight", "heads.aff.blocks.0.conv1.weight", "heads.aff.input_projection.weight",
              "heads.sdt.projection.bias", "model.stem.weight", "model.enc_block_0.0.conv3.weight", "model.up_0.res_conv.weight")


def main():
    M = R.ref_mednext_models()
    net = M.build_mednext_custom(make_cfg()).train()
    assert type(net).__name__ == "MedNeXtMultiHeadWrapper"
    fill_deterministic(net)
    x, g = inputs()
    out = net(x)["output"]
    sum((out[k] * g[k]).sum() for k in HEADS).backward()
    arrays = {f"out_{k}": out[k].detach().numpy() for k in HEADS}
    names, norms, dots = [], [], []
    for name, p in net.named_parameters():
        if p.grad is None:
            continue
        names.append(name)
        norms.append(float(p.grad.norm()))
        dots.append(float((p.grad * probe(name, p.shape)).sum()))
        if name in FULL_GRADS:
            arrays[f"grad::{name}"] = p.grad.numpy()
    arrays["grad_names"] = np.array(names)
    arrays["grad_norms"] = np.array(norms, dtype=np.float64)
    arrays["grad_dots"] = np.array(dots, dtype=np.float64)
    missing = [n for n in FULL_GRADS if f"grad::{n}" not in arrays]
    assert not missing, missing
    np.savez_compressed(OUT, **arrays)
    print(f"wrote {OUT}: {len(names)} gradient records, {len(FULL_GRADS)} full gradients, {os.path.getsize(OUT) / 1e6:.2f} MB; "
          f"|out| aff {float(out["aff"].detach().abs().mean()):.3f} sdt {float(out["sdt"].detach().abs().mean()):.3f}")


if __name__ == "__main__":
    main()

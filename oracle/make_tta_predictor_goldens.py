"""TEST INFRASTRUCTURE — golden vectors from the REAL ``connectomics/inference/tta.py::TTAPredictor`` (executed in place with
its real dependencies and the real ``EagerSlidingWindowEngine`` of ``window.py``): volume-first and patch-first-local TTA
through a sliding-window engine, with activations, channel selection, per-channel ensemble modes, a mask, and the
disabled-TTA path.  fp32 on the CPU.  Build-container only; writes ``tests/golden/tta_predictor_goldens.npz``.
Run: ``python -m oracle.make_tta_predictor_goldens``."""

from __future__ import annotations

import os
from types import SimpleNamespace as NS

import numpy as np
import torch

from . import ref_loader as R
from . import tta_oracle as TO

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tta_predictor_goldens.npz")

ACTS = [dict(channels="0:2", activation="sigmoid"), dict(channels="2:3", activation="tanh")]
ENGINE = dict(roi_size=(16, 16, 16), sw_batch_size=2, overlap=0.5, mode="bump", padding_mode="constant", cval=0.0)

CASES = {
    "volume_first": dict(tta=dict(enabled=True, flip_axes="all", rotation90_axes=None, rotate90_k=None, ensemble_mode="mean",
                                  patch_first_local=False), acts=ACTS, select=[2, 0], engine=True, mask=True),
    "patch_first": dict(tta=dict(enabled=True, flip_axes="all", rotation90_axes=[[1, 2]], rotate90_k=[0, 1],
                                 ensemble_mode=[["0:1", "min"], ["1:3", "mean"]], patch_first_local=True), acts=ACTS, select=None,
                        engine=True, mask=False),
    "direct_rot": dict(tta=dict(enabled=True, flip_axes=[[0], [1, 2]], rotation90_axes=[[1, 2]], rotate90_k=[0, 1, 3],
                                ensemble_mode="max", patch_first_local=False), acts=[dict(channels=[0, 1], activation="softmax")],
                       select="0:2", engine=False, mask=True),
    "disabled": dict(tta=dict(enabled=False), acts=ACTS, select=None, engine=True, mask=False),
}


def make_cfg(case: dict):
    """the ONE place a golden case becomes a config object (the test imports it)"""
    sw = NS(window_size=list(ENGINE["roi_size"]), overlap=ENGINE["overlap"], sw_batch_size=ENGINE["sw_batch_size"], blending=ENGINE["mode"],
            padding_mode=ENGINE["padding_mode"], cval=ENGINE["cval"], keep_input_on_cpu=False, sw_device=None, output_device=None)
    return NS(model=NS(out_channels=3, primary_head=None, heads=None, output_size=list(ENGINE["roi_size"])),
              data=NS(dataloader=NS(batch_size=2)),
              inference=NS(test_time_augmentation=NS(apply_mask=True, distributed_sharding=False, **case["tta"]), sliding_window=sw,
                           model=NS(channel_activations=case["acts"], select_channel=case["select"], output_dtype=None, head=None)))


def inputs():
    rs = np.random.RandomState(77)
    x = torch.from_numpy(rs.rand(1, 1, 24, 32, 32).astype(np.float32))
    mask = torch.from_numpy((rs.rand(24, 32, 32) > 0.35).astype(np.float32))
    return x, mask


def main():
    T = R.ref_tta()
    W = R.ref_window()
    x, mask = inputs()
    net = TO.ramp_network(3)
    arrays = {}
    for name, case in CASES.items():
        cfg = make_cfg(case)
        engine = W.EagerSlidingWindowEngine(sw_device=None, output_device=None, **ENGINE) if case["engine"] else None
        pred = T.TTAPredictor(cfg, engine, net)
        out = pred.predict(x.clone(), mask=mask.clone() if case["mask"] else None)
        arrays[name] = out.numpy()
        print(f"{name}: {tuple(out.shape)} mean {float(out.mean()):.4f} activation types {pred.channel_activation_types}")
    np.savez_compressed(OUT, **arrays)
    print(f"wrote {OUT} ({os.path.getsize(OUT) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()

"""Generate tests/golden/*.npz by running the REAL reference files (build container only).

    python -m oracle.make_goldens

The reference's sliding-window module is executed in place from /root/reference (see
ref_loader.py); its outputs become the committed golden vectors that pin both the oracle
restatement (tests/test_oracle_window.py, CPU) and the CUDA path (tests/test_sw_gpu.py, GPU).
The network arithmetic lives in the un-vendored `nnunet_mednext`, so the MedNeXt golden is
produced by the oracle restatement itself (a regression vector, flagged `source=oracle`).
"""

from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_loader
from .mednext_oracle import MedNeXt

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

GRID_CASES = [
    # (image, roi, overlap)
    ((24, 24, 24), (8, 8, 8), 0.5),
    ((20, 23, 17), (8, 8, 8), 0.5),
    ((20, 23, 17), (8, 6, 5), (0.5, 0.25, 0.0)),
    ((5, 30, 9), (8, 8, 8), 0.5),
    ((165, 1024, 768), (112, 112, 112), 0.5),
    ((2048, 2048, 2048), (160, 160, 160), 0.5),
    ((320, 320, 320), (160, 160, 160), 0.5),
    ((100, 100, 100), (32, 32, 32), 0.3),
    ((33, 33, 33), (32, 32, 32), 0.99),
    ((64, 64, 64), (10, 10, 10), 0.75),   # roi*(1-ov)=2.5 -> banker's rounding -> 2
]


def affine_net(x):
    """Deterministic 2-channel stand-in network: [2x+1, -x+0.5]."""
    return torch.cat([2.0 * x[:, :1] + 1.0, -1.0 * x[:, :1] + 0.5], dim=1)


def main():
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    W = ref_loader.ref_window()
    g = {}

    # ---- integer grid logic
    for i, (img, roi, ov) in enumerate(GRID_CASES):
        iv = W.compute_scan_interval(img, roi, overlap=ov)
        g[f"grid{i}_interval"] = np.asarray(iv, dtype=np.int64)
        starts = W.dense_patch_slices(img, roi, iv, return_slice=False)
        g[f"grid{i}_count"] = np.asarray([len(starts)], dtype=np.int64)
        if len(starts) <= 20000:
            g[f"grid{i}_starts"] = np.asarray(starts, dtype=np.int64)
        # per-axis starts (always small)
        for a in range(3):
            ax = sorted({s[a] for s in starts})
            g[f"grid{i}_axis{a}"] = np.asarray(ax, dtype=np.int64)

    # ---- importance maps
    for name, roi in [("r8", (8,)), ("r675", (6, 7, 5)), ("r16", (16, 16, 16)), ("r444", (4, 4, 4))]:
        for mode in ("bump", "constant", "distance_transform"):
            for dt, dn in ((torch.float32, "f32"), (torch.float16, "f16")):
                m = W.build_sliding_importance_map(roi, mode=mode, device="cpu", dtype=dt)
                g[f"imap_{name}_{mode}_{dn}"] = m.float().numpy()
    m160 = W.build_sliding_importance_map((160, 160, 160), mode="bump", device="cpu", dtype=torch.float32)
    g["imap_160_probe"] = np.asarray([m160[80, 80, 80], m160[79, 79, 79], m160[0, 0, 0], m160[10, 80, 80],
                                      m160.double().sum()], dtype=np.float64)
    g["imap_160_line"] = m160[:, 80, 80].numpy()

    # ---- normaliser
    torch.manual_seed(3)
    v = torch.randn(1, 2, 5, 6, 7)
    w = torch.rand(1, 1, 5, 6, 7) * 2e-4
    g["norm_in_v"] = v.numpy().copy()
    g["norm_in_w"] = w.numpy().copy()
    g["norm_out_f32"] = W.normalize_weighted_accumulator(v.clone(), w.clone()).numpy()
    g["norm_out_f16"] = W.normalize_weighted_accumulator(v.clone().half(), w.clone().half()).float().numpy()

    # ---- padded patch extraction
    torch.manual_seed(4)
    vol = torch.randn(1, 2, 9, 10, 11)
    g["patch_vol"] = vol.numpy()
    for k, (st, roi, mode) in enumerate([((-3, 2, 5), (8, 8, 8), "constant"), ((4, 6, 7), (8, 8, 8), "reflect"),
                                         ((-2, -2, -2), (6, 6, 6), "replicate"), ((5, 0, 0), (8, 8, 8), "reflect")]):
        sl = [tuple(slice(s, s + r) for s, r in zip(st, roi))]
        p, _ = W._extract_padded_patch_batch(vol, sl, roi_size=roi, padding_mode=mode, cval=0.25)
        g[f"patch{k}"] = p.numpy()
        g[f"patch{k}_meta"] = np.asarray(list(st) + list(roi), dtype=np.int64)

    # ---- eager engine end-to-end
    def engine(roi, ov, mode, pad, bs=2, cval=0.0):
        return W.EagerSlidingWindowEngine(roi_size=roi, sw_batch_size=bs, overlap=ov, mode=mode,
                                          padding_mode=pad, cval=cval, sw_device=None, output_device=None)

    ar = torch.arange(24 ** 3, dtype=torch.float32).view(1, 1, 24, 24, 24) / 1000.0
    g["eng_identity_const"] = engine((8, 8, 8), 0.5, "constant", "constant")(inputs=ar, network=lambda x: x).numpy()
    g["eng_identity_bump"] = engine((8, 8, 8), 0.5, "bump", "constant")(inputs=ar, network=lambda x: x).numpy()
    torch.manual_seed(5)
    x = torch.rand(1, 1, 20, 23, 17)
    g["eng_in"] = x.numpy()
    g["eng_affine_bump"] = engine((8, 8, 8), 0.5, "bump", "constant", bs=3)(inputs=x, network=affine_net).numpy()
    g["eng_affine_dt"] = engine((8, 6, 5), (0.5, 0.25, 0.0), "distance_transform", "constant")(inputs=x, network=affine_net).numpy()
    g["eng_affine_reflect"] = engine((8, 8, 8), 0.25, "bump", "reflect")(inputs=x, network=affine_net).numpy()
    small = torch.rand(1, 1, 5, 30, 9)
    g["eng_small_in"] = small.numpy()
    g["eng_small_bump"] = engine((8, 8, 8), 0.5, "bump", "constant", cval=0.5)(inputs=small, network=affine_net).numpy()
    g["eng_affine_bump_f16"] = engine((8, 8, 8), 0.5, "bump", "constant")(
        inputs=x, network=lambda t: affine_net(t).half()).float().numpy()

    # ---- chunk grid / halo
    CG = ref_loader.ref_chunk_grid()
    H = ref_loader.ref_halo()
    chunks = CG.build_chunk_grid((100, 64, 70), (48, 64, 32))
    g["chunk_starts"] = np.asarray([c.start for c in chunks], dtype=np.int64)
    g["chunk_stops"] = np.asarray([c.stop for c in chunks], dtype=np.int64)
    halos = [H.resolve_halo_region(c, (100, 64, 70), halo=(8, 4, 6)) for c in chunks]
    g["halo_read_start"] = np.asarray([h[0] for h in halos], dtype=np.int64)
    g["halo_read_stop"] = np.asarray([h[1] for h in halos], dtype=np.int64)
    g["halo_core"] = np.asarray([[s.start for s in h[2]] + [s.stop for s in h[2]] for h in halos], dtype=np.int64)

    np.savez_compressed(os.path.join(OUT, "window_goldens.npz"), **g)
    print("wrote window_goldens.npz", len(g), "arrays")

    # ---- MedNeXt regression vector from the oracle restatement (source=oracle)
    torch.manual_seed(0)
    net = MedNeXt(in_channels=1, n_channels=16, n_classes=2, exp_r=2, kernel_size=3,
                  deep_supervision=True, do_res=True, do_res_up_down=True, block_counts=[1] * 9).eval()
    torch.manual_seed(1)
    xin = torch.rand(1, 1, 32, 32, 32)
    with torch.no_grad():
        outs = net(xin)
    np.savez_compressed(os.path.join(OUT, "mednext_tiny.npz"), x=xin.numpy(),
                        **{f"out{i}": o.numpy() for i, o in enumerate(outs)},
                        n_params=np.asarray([sum(p.numel() for p in net.parameters())]))
    print("wrote mednext_tiny.npz")


if __name__ == "__main__":
    main()

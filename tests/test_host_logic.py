"""CPU-side tests: the C-ABI library loads and exports every declared symbol; integer grid logic
(host C code) is bit-exact with the reference-generated golden vectors; registry seam semantics
(reference tests/unit/test_architecture_registry.py:55-200)."""
import ctypes
import os
import warnings
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle.make_goldens import GRID_CASES
from oracle import window_oracle as O
from pytorch_connectomics_b200 import _lib
from pytorch_connectomics_b200 import architectures as A
from pytorch_connectomics_b200.inference import window as W


@pytest.fixture(scope="module", autouse=True)
def _built():
    _lib.build()


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _lib.exported_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert lib.pcb_version() >= 100


def test_grid_bit_exact_vs_reference_goldens(window_goldens):
    g = window_goldens
    for i, (img, roi, ov) in enumerate(GRID_CASES):
        assert list(W.compute_scan_interval(img, roi, overlap=ov)) == g[f"grid{i}_interval"].tolist()
        starts = W._plan(_lib.GRID_EAGER, img, roi, ov)
        assert len(starts) == int(g[f"grid{i}_count"][0])
        if f"grid{i}_starts" in g:
            assert np.array_equal(np.asarray(starts, dtype=np.int64), g[f"grid{i}_starts"])
            iv = W.compute_scan_interval(img, roi, overlap=ov)
            assert W.dense_patch_slices(img, roi, iv, return_slice=False) == starts
        for a in range(3):
            assert sorted({s[a] for s in starts}) == g[f"grid{i}_axis{a}"].tolist()


def test_lazy_grid_matches_oracle():
    for img, roi, ov in [((2048,) * 3, (160,) * 3, 0.5), ((12, 14, 13), (6, 6, 6), 0.5), ((100, 90, 80), (32, 16, 24), 0.3)]:
        for snap, kind in ((False, _lib.GRID_LAZY), (True, _lib.GRID_LAZY_SNAP)):
            per_axis = O.lazy_axis_offsets(img, roi, (ov,) * 3, snap)
            n = len(per_axis[0]) * len(per_axis[1]) * len(per_axis[2])
            starts = W._plan(kind, img, roi, ov)
            assert len(starts) == n
            if n < 50000:
                for a in range(3):
                    assert sorted({s[a] for s in starts}) == per_axis[a]
    recs = O.lazy_region_records((12, 14, 13), (6, 6, 6), (0.5,) * 3, (3, 2, 4), (9, 11, 13), False)
    starts = W._plan(_lib.GRID_LAZY, (12, 14, 13), (6, 6, 6), 0.5, region=((3, 2, 4), (9, 11, 13)))
    assert [r[0] for r in recs] == starts


def test_scan_interval_errors_and_rounding():
    assert W.compute_scan_interval((64,) * 3, (10,) * 3, overlap=0.75) == (2, 2, 2)    # 2.5 -> 2 (half-even)
    assert W.compute_scan_interval((64,) * 3, (14,) * 3, overlap=0.75) == (4, 4, 4)    # 3.5 -> 4
    assert W.compute_scan_interval((5, 64), (8, 8), overlap=(0.5, 2.0)) == (5, 1)       # clamp 0.99 -> max(1, round(.08))
    with pytest.raises(ValueError):
        W.compute_scan_interval((8, 8, 8), (0, 8, 8), overlap=0.5)


def test_registry_seam():
    assert A.is_architecture_available("mednext") and A.is_architecture_available("mednext_custom")
    assert A.list_architectures() == sorted(A.list_architectures())
    with pytest.raises(ValueError, match="not found"):
        A.get_architecture_builder("nope")

    @A.register_architecture("tmp_arch")
    def _b(cfg):
        """doc line"""
        return torch.nn.Identity()

    assert A.get_architecture_builder("tmp_arch") is _b
    assert A.get_architecture_info()["tmp_arch"]["doc"] == "doc line"
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        A.register_architecture("tmp_arch")(_b)
    assert any(issubclass(w.category, UserWarning) for w in rec)
    A.unregister_architecture("tmp_arch")
    with pytest.raises(ValueError):
        A.unregister_architecture("tmp_arch")


def _cfg(**mednext):
    return NS(model=NS(arch=NS(type="mednext"), in_channels=1, out_channels=2, mednext=NS(**mednext),
                       loss=NS(deep_supervision=False)))


def test_builders_validation_and_state_dict_parity():
    from oracle.mednext_oracle import create_mednext_v1
    m = A.build_model(_cfg(size="S", kernel_size=3))
    info = m.get_model_info()
    assert info["name"] == "MedNeXtWrapper" and info["parameters"] == info["trainable_parameters"]
    ref = create_mednext_v1(1, 2, "S", 3, False)
    assert list(ref.state_dict().keys()) == list(m.model.state_dict().keys())
    m.model.load_state_dict(ref.state_dict(), strict=True)
    with pytest.raises(ValueError, match="model_size"):
        A.build_model(_cfg(size="XL", kernel_size=3))
    with pytest.raises(ValueError, match="kernel_size"):
        A.build_model(_cfg(size="S", kernel_size=4))
    with pytest.raises(ValueError, match="checkpoint_style"):
        A.build_model(_cfg(size="S", kernel_size=3, checkpoint_style="inside"))
    assert A.build_model(_cfg(size="S", kernel_size=3, checkpoint_style="outside_block")).model.outside_block_checkpointing
    cfg = _cfg(size="S", kernel_size=3)
    cfg.model.heads = {"aff": {"out_channels": 3, "num_blocks": 1}, "sdt": {"out_channels": 1}}
    mh = A.build_model(cfg)
    assert mh.primary_head == "aff" and mh.feature_channels == 32
    assert mh.head_specs["sdt"] == {"out_channels": 1, "num_blocks": 0, "hidden_channels": 32}
    cfg.model.primary_head = "nope"
    with pytest.raises(ValueError, match="primary_head"):
        A.build_model(cfg)


def test_no_cpu_fallback():
    m = A.build_model(_cfg(size="S", kernel_size=3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 32, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        W.build_sliding_importance_map((8, 8, 8), mode="bump", device="cpu")
    eng = W.EagerSlidingWindowEngine(roi_size=(8, 8, 8), sw_batch_size=1, overlap=0.5, mode="bump",
                                     padding_mode="constant", cval=0.0)
    with pytest.raises(ValueError):
        eng(inputs=torch.zeros(2, 1, 8, 8, 8), network=lambda t: t)
    with pytest.raises(ValueError):
        eng(inputs=torch.zeros(1, 8, 8), network=lambda t: t)
    # the newer seams fail just as loudly: TTA views / folds, the z-slab engine, LayerNorm blocks
    from pytorch_connectomics_b200.inference import TTAEnsemble, ZSlabShardedEngine, apply_view
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        apply_view(torch.zeros(1, 1, 4, 4, 4), [0], None, 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TTAEnsemble(NS(enabled=True, flip_axes="all")).predict(torch.zeros(1, 1, 4, 4, 4), lambda t: t)
    with pytest.raises(ValueError, match=r"\[N,C,D,H,W\]"):
        TTAEnsemble(None).predict(torch.zeros(1, 4, 4, 4), lambda t: t)
    sh = ZSlabShardedEngine(roi_size=(8, 8, 8), sw_batch_size=1, overlap=0.5, mode="bump", rank=0, world=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        sh(torch.zeros(1, 1, 16, 16, 16), lambda t: t)
    with pytest.raises(ValueError, match="smaller than the window"):
        sh(torch.zeros(1, 1, 4, 16, 16), lambda t: t)
    with pytest.raises(ValueError, match="batch size 1"):
        sh(torch.zeros(2, 1, 16, 16, 16), lambda t: t)
    with pytest.raises(ValueError, match="3-D roi_size"):
        ZSlabShardedEngine(roi_size=(8, 8), sw_batch_size=1, overlap=0.5, mode="bump")
    from pytorch_connectomics_b200.architectures.mednext import MedNeXtBlock
    blk = MedNeXtBlock(16, 16, 2, 3, norm_type="layer")
    assert type(blk.norm).__name__ == "LayerNorm" and list(blk.state_dict()) == list(MedNeXtBlock(16, 16, 2, 3).state_dict())
    with pytest.raises(ValueError, match="norm_type"):
        MedNeXtBlock(16, 16, 2, 3, norm_type="batch")
    with pytest.raises(NotImplementedError):
        MedNeXtBlock(16, 16, 2, 3, n_groups=4)
    assert tuple(MedNeXtBlock(16, 16, 2, 3, grn=True).grn_gamma.shape) == (1, 32, 1, 1, 1)      # upstream's GRN parameters
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk(torch.zeros(1, 4, 4, 4, 16, dtype=torch.bfloat16))


def test_config_resolvers():
    cfg = NS(inference=NS(sliding_window=NS(window_size=[16, 32, 32], overlap=[0.5, 1.5, -1], sw_batch_size=None,
                                            blending="Bump ", padding_mode="reflect", cval=0.5, border_mask=[2])),
             data=NS(dataloader=NS(batch_size=3)))
    assert W.resolve_inferer_roi_size(cfg) == (16, 32, 32)
    assert W.resolve_inferer_overlap(cfg, (16, 32, 32)) == (0.5, 0.99, 0.0)
    rt = W._resolve_sliding_window_runtime(cfg, (16, 32, 32))
    assert rt["sw_batch_size"] == 3 and rt["mode"] == "bump" and rt["padding_mode"] == "reflect" and rt["cval"] == 0.5
    assert W.resolve_border_mask(cfg, 3) == [2, 2, 2]
    eng = W.build_sliding_inferer(cfg)
    assert eng.roi_size == (16, 32, 32) and eng.sw_batch_size == 3
    assert W.build_sliding_inferer(NS()) is None
    assert W.resolve_model_output_dtype(NS(inference=NS(model=NS(output_dtype="torch.float16")))) == torch.float16
    with pytest.raises(ValueError):
        W.resolve_model_output_dtype(NS(inference=NS(model=NS(output_dtype="int8"))))
    assert W.is_distance_transform_blending("BANIS")


def test_chunk_grid_and_halo_vs_reference_goldens(window_goldens):
    from pytorch_connectomics_b200.inference import chunked as C
    g = window_goldens
    chunks = C.build_chunk_grid((100, 64, 70), (48, 64, 32))
    assert np.array_equal(np.asarray([c.start for c in chunks]), g["chunk_starts"])
    assert np.array_equal(np.asarray([c.stop for c in chunks]), g["chunk_stops"])
    assert chunks[0].key == "z0_y0_x0" and chunks[-1].key == "z2_y0_x2"
    halos = [C.resolve_halo_region(c, (100, 64, 70), halo=(8, 4, 6)) for c in chunks]
    assert np.array_equal(np.asarray([h[0] for h in halos]), g["halo_read_start"])
    assert np.array_equal(np.asarray([h[1] for h in halos]), g["halo_read_stop"])
    assert np.array_equal(np.asarray([[s.start for s in h[2]] + [s.stop for s in h[2]] for h in halos]), g["halo_core"])
    # round-robin rank assignment (chunked.py:471) partitions the grid
    owned = [i for r in range(4) for i, _ in C.chunks_for_rank(chunks, r, 4)]
    assert sorted(owned) == list(range(len(chunks)))
    assert [i for i, _ in C.chunks_for_rank(chunks, 1, 4)] == [1, 5]
    with pytest.raises(ValueError):
        C.build_chunk_grid((10, 10), (4, 4, 4))
    assert C.resolve_chunk_shape((32, 32, 32), (100, 20, 64), "z") == (32, 20, 64)
    assert C.resolve_chunk_shape((32, 32, 32), (100, 20, 64)) == (32, 20, 32)
    assert C.resolve_external_chunk_shard(None, None) is None and C.resolve_external_chunk_shard(1, 4) == (1, 4)
    for bad in ((None, 2), (2, 2), (0, 0)):
        with pytest.raises(ValueError):
            C.resolve_external_chunk_shard(*bad)


def test_lazy_records_match_oracle():
    from pytorch_connectomics_b200.inference.lazy import lazy_window_records
    for snap in (False, True):
        got = lazy_window_records((12, 14, 13), (6, 6, 6), (0.5,) * 3, (3, 2, 4), (9, 11, 13), snap)
        want = O.lazy_region_records((12, 14, 13), (6, 6, 6), (0.5,) * 3, (3, 2, 4), (9, 11, 13), snap)
        assert [(r[0], r[1], tuple(h - l for l, h in zip(r[1], r[2])), r[3]) for r in want] == got


def _ddp_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorch_connectomics_b200.training import FlatGradArena, broadcast_parameters
    torch.manual_seed(rank)                # ranks start from DIFFERENT weights ...
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.ReLU(), torch.nn.Linear(8, 2), torch.nn.Linear(2, 2),
                              torch.nn.BatchNorm1d(2))
    net[4].running_mean.fill_(float(rank))
    broadcast_parameters(net, src=0)       # ... and rank 0's parameters AND buffers win (DDP construction semantics)
    arena = FlatGradArena(net.parameters())
    x = torch.full((3, 4), float(rank + 1))
    if rank == 1:                          # Lightning-style zero_grad detaches the views: the arena must repair them
        net.zero_grad(set_to_none=True)
    else:
        arena.zero()
    net[2](net[1](net[0](x))).sum().backward()
    arena.allreduce()
    assert all(p.grad is not None and p.grad.data_ptr() >= arena.buffer.data_ptr() for p in net.parameters())
    q.put((rank, arena.packed().numpy(), torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy(),
           float(net[4].running_mean[0])))   # by value: torch tensors travel as fds that die with the worker
    dist.destroy_process_group()


def test_flat_arena_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    from conftest import free_port
    port = free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    res = {r: torch.from_numpy(v) for r, v, _, _ in got}
    par = {r: torch.from_numpy(w) for r, _, w, _ in got}
    for p in procs:
        p.join(timeout=60)
    assert torch.equal(par[0], par[1]) and all(rm == 0.0 for _, _, _, rm in got)   # broadcast: params and buffers of rank 0
    assert torch.equal(res[0], res[1])                      # identical averaged gradients on both ranks
    # reference: mean of the two per-rank gradients computed in-process
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.ReLU(), torch.nn.Linear(8, 2), torch.nn.Linear(2, 2),
                              torch.nn.BatchNorm1d(2))
    gs = []
    for r in range(2):
        net.zero_grad()
        net[2](net[1](net[0](torch.full((3, 4), float(r + 1))))).sum().backward()
        gs.append(torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in net.parameters()]))
    assert torch.allclose(res[0], (gs[0] + gs[1]) / 2, atol=1e-6)
    assert res[0][-10:].abs().sum() == 0                    # the unused head + norm stay zero (no hang, no None grads)


def _reduce_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorch_connectomics_b200.inference.lazy_distributed import (make_accumulator_reducer, should_shard_windows,
                                                                      validate_patch_shard)
    assert should_shard_windows(True) and not should_shard_windows(False)
    validate_patch_shard(3, 6, "cpu")
    v = torch.full((1, 2, 3, 3, 3), float(rank + 1))
    w = torch.full((1, 1, 3, 3, 3), 0.5)
    out = make_accumulator_reducer()(v, w)
    q.put((rank, None if out is None else (out[0].clone().numpy(), out[1].clone().numpy())))
    try:
        validate_patch_shard(0 if rank == 1 else 2, 2, "cpu")
        q.put((rank, "no error"))
    except RuntimeError as e:
        q.put((rank, "empty" in str(e)))
    # the reference's own names and signatures (lazy_distributed.py:10-169)
    from types import SimpleNamespace as NS
    from pytorch_connectomics_b200.inference import lazy_distributed as LD
    assert LD.distributed_context() == (True, rank, world)
    cfg = NS(inference=NS(sliding_window=NS(distributed_sharding=True)), data=NS(dataloader=NS(use_lazy_zarr=True)))
    assert LD.is_distributed_window_sharding_enabled(cfg)
    cfg.data.dataloader.use_lazy_zarr = False
    assert not LD.is_distributed_window_sharding_enabled(cfg) and not LD.is_distributed_window_sharding_enabled(NS())
    assert LD.distributed_reduction_device(torch.device("cpu")).type in ("cpu", "cuda")
    cpu = torch.device("cpu")
    LD.validate_distributed_tensor_shape(torch.zeros(2, 3), name="acc", reduction_device=cpu)
    try:
        LD.validate_distributed_tensor_shape(torch.zeros(2, 3 + rank), name="acc", reduction_device=cpu)
        shape_err = "no error"
    except RuntimeError as e:
        shape_err = str(e)
    # a strided 1.5 MB accumulator takes the staged route in two 1 MB pieces
    big = torch.full((2, 200_000), float(rank + 1)).t()
    red = LD.reduce_cpu_tensor_to_rank_zero(big, op=dist.ReduceOp.SUM, reduction_device=cpu, chunk_mb=1, name="value accumulator")
    hook = LD.make_accumulator_reduce_hook(reduction_device=cpu, chunk_mb=1)
    pair = hook(torch.full((4,), float(rank)), torch.ones(4))
    # the REAL lazy_distributed.py executed in place gives the same results and the same error texts (build container only)
    from oracle import ref_loader
    if ref_loader.available():
        ref_loader._base_stubs()
        RD = ref_loader._load("connectomics.inference.lazy_distributed", "connectomics/inference/lazy_distributed.py")
        assert RD.distributed_context() == LD.distributed_context()
        cfg.data.dataloader.use_lazy_zarr = True
        assert RD.is_distributed_window_sharding_enabled(cfg) == LD.is_distributed_window_sharding_enabled(cfg) is True
        assert RD.distributed_reduction_device(cpu) == LD.distributed_reduction_device(cpu)

        def outcome(fn):
            try:
                return ("ok", fn())
            except RuntimeError as e:
                return ("RuntimeError", str(e))

        for mod_pair in [(lambda M: M.validate_distributed_tensor_shape(torch.zeros(2, 3 + rank), name="acc", reduction_device=cpu)),
                         (lambda M: M.validate_distributed_tensor_shape(torch.zeros(*([1] * 9)), name="acc", reduction_device=cpu)),
                         (lambda M: M.validate_distributed_patch_shard(local_count=rank, total_count=1, reduction_device=cpu)),
                         (lambda M: M.validate_distributed_patch_shard(local_count=2, total_count=4, reduction_device=cpu))]:
            assert outcome(lambda: mod_pair(RD)) == outcome(lambda: mod_pair(LD))
        big2 = torch.full((2, 200_000), float(rank + 1)).t()
        a = RD.reduce_cpu_tensor_to_rank_zero(big2.clone(), op=dist.ReduceOp.SUM, reduction_device=cpu, chunk_mb=1, name="value accumulator")
        b = LD.reduce_cpu_tensor_to_rank_zero(big2.clone(), op=dist.ReduceOp.SUM, reduction_device=cpu, chunk_mb=1, name="value accumulator")
        assert (a is None) == (b is None) == (rank != 0) and (a is None or torch.equal(a, b))
        pa = RD.make_accumulator_reduce_hook(reduction_device=cpu, chunk_mb=1)(torch.full((4,), float(rank)), torch.ones(4))
        pb = LD.make_accumulator_reduce_hook(reduction_device=cpu, chunk_mb=1)(torch.full((4,), float(rank)), torch.ones(4))
        assert (pa is None) == (pb is None) and (pa is None or (torch.equal(pa[0], pb[0]) and torch.equal(pa[1], pb[1])))
    q.put((rank, ("ref-names", shape_err, None if red is None else (tuple(red.shape), float(red.min()), float(red.max())),
                  None if pair is None else (pair[0].tolist(), pair[1].tolist()))))
    dist.destroy_process_group()


def test_accumulator_reduce_to_root_gloo_world2():
    # reference lazy_distributed.py:78-169: SUM onto rank 0, None elsewhere, empty-shard check
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    from conftest import free_port
    port = free_port()
    procs = [ctx.Process(target=_reduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(6)]
    for p in procs:
        p.join(timeout=60)
    is_named = lambda v: isinstance(v, tuple) and len(v) == 4 and isinstance(v[0], str) and v[0] == "ref-names"
    named = {r: v for r, v in got if is_named(v)}
    got = [(r, v) for r, v in got if not is_named(v)]
    for r in range(2):
        _, shape_err, red, pair = named[r]
        assert "same shape" in shape_err and "rank 0: (2, 3)" in shape_err and "rank 1: (2, 4)" in shape_err
        assert red == (((200_000, 2), 3.0, 3.0) if r == 0 else None)
        assert pair == (([1.0] * 4, [2.0] * 4) if r == 0 else None)
    first = {r: v for r, v in got if not isinstance(v, (bool, str))}
    assert first[1] is None
    assert torch.equal(torch.from_numpy(first[0][0]), torch.full((1, 2, 3, 3, 3), 3.0))
    assert torch.equal(torch.from_numpy(first[0][1]), torch.full((1, 1, 3, 3, 3), 1.0))
    assert all(v is True for r, v in got if isinstance(v, (bool, str)))


def test_monai_unet_builder_and_state_dict_keys():
    from oracle.monai_unet_oracle import UNet as OracleUNet
    cfg = NS(model=NS(arch=NS(type="monai_unet"), in_channels=1, out_channels=1, input_size=[32, 64, 64],
                      monai=NS(filters=[16, 32, 64], num_res_units=1, kernel_size=3, dropout=0.0)))
    m = A.build_model(cfg)
    assert type(m).__name__ == "MONAIModelWrapper" and not m.supports_deep_supervision
    ref = OracleUNet(3, 1, 1, [16, 32, 64], [2, 2], num_res_units=1)
    sd_ref, sd = ref.state_dict(), m.model.state_dict()
    assert list(sd_ref.keys()) == list(sd.keys())
    assert all(sd_ref[k].shape == sd[k].shape for k in sd_ref)
    m.model.load_state_dict(sd_ref, strict=True)
    assert "model.0.conv.unit0.adn.A.weight" in sd and "model.1.submodule.2.0.conv.weight" in sd
    assert m.get_model_info()["parameters"] == sum(p.numel() for p in ref.parameters())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 32, 64, 64))
    cfg.model.monai.norm = "group"            # 8 groups by default: torch refuses GroupNorm(8, 1) at the 1-channel output, as MONAI does
    with pytest.raises(ValueError, match="divisible by num_groups"):
        A.build_model(cfg)
    cfg.model.monai.norm, cfg.model.monai.upsample_mode = "batch", "nontrainable"
    assert "model.2.0.preconv.weight" in A.build_model(cfg).model.state_dict()        # UpsampleModeUNet's interpolating up path
    cfg.model.monai.upsample_mode = "pixelshuffle"
    with pytest.raises(NotImplementedError):
        A.build_model(cfg)


def test_apply_border_mask_matches_reference():
    # window.py:297-319: zero the outer k voxels per axis; too-large masks raise
    from oracle import ref_loader
    m = W.apply_border_mask(torch.ones(6, 8, 10), [1, 2, 0])
    assert m.sum() == (6 - 2) * (8 - 4) * 10 and m[0].sum() == 0 and m[:, :2].sum() == 0 and m[:, :, 0].sum() > 0
    assert torch.equal(W.apply_border_mask(torch.ones(4, 4), []), torch.ones(4, 4))
    with pytest.raises(ValueError, match="too large"):
        W.apply_border_mask(torch.ones(4, 4, 4), [2, 0, 0])
    if ref_loader.available():      # build container: the real reference file, executed in place
        R = ref_loader.ref_window()
        for shape, mask in [((6, 8, 10), [1, 2, 0]), ((5, 5, 5), [2, 1, 1]), ((2, 7, 9), [0, 3, 4])]:
            torch.manual_seed(0)
            a = torch.rand(*shape)
            assert torch.equal(W.apply_border_mask(a.clone(), mask), R.apply_border_mask(a.clone(), mask))


def test_upkern_load_weights_resizes_depthwise_kernels():
    """mednext_models.py:487-537 (UpKern): equal-shaped tensors are copied, k=3 depthwise / transposed kernels are resized to
    k=5 by trilinear interpolation — checked against F.interpolate on the same tensors, through wrappers and bare modules."""
    import torch.nn.functional as F
    from pytorch_connectomics_b200.architectures import mednext as PM
    torch.manual_seed(0)
    kw = dict(in_channels=1, n_channels=16, n_classes=2, exp_r=2, deep_supervision=False, do_res=True, do_res_up_down=True,
              block_counts=[1] * 9)
    small, big = PM.MedNeXt(kernel_size=3, **kw), PM.MedNeXt(kernel_size=5, **kw)
    before = {k: v.clone() for k, v in big.state_dict().items()}
    out = PM.upkern_load_weights(PM.MedNeXtWrapper(big, deep_supervision=False), PM.MedNeXtWrapper(small, deep_supervision=False))
    assert out.model is big
    ssd, bsd = small.state_dict(), big.state_dict()
    resized = 0
    for k, v in bsd.items():
        if ssd[k].shape == v.shape:
            assert torch.equal(v, ssd[k]), k
        else:
            assert v.shape[2:] == (5, 5, 5) and ssd[k].shape[2:] == (3, 3, 3)
            assert torch.allclose(v, F.interpolate(ssd[k], size=(5, 5, 5), mode="trilinear")), k
            assert not torch.equal(v, before[k])
            resized += 1
    assert resized == 9 + 4 + 4            # conv1 of 9 stages' blocks, 4 down and 4 up blocks
    bad = PM.MedNeXt(kernel_size=5, **{**kw, "n_channels": 32})
    with pytest.raises(ValueError, match="identical architecture"):
        PM.upkern_load_weights(bad, small)


def test_param_groups_equal_the_real_build_optimizer():
    """`reference_param_groups` / `build_fused_adamw`'s grouping against the REAL `training/optimization/build.py::build_optimizer`
    (executed in place) on this package's own module trees: same parameter order, same lr / weight_decay per parameter —
    MedNeXt with GroupNorm and with the channels-first LayerNorm (a plain nn.Module upstream: NOT a norm layer for the
    reference's rule), the multi-head wrapper, the MONAI U-Net (BatchNorm + PReLU), shared parameters counted once."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference is only present in the build container")
    from pytorch_connectomics_b200.architectures import mednext as PM
    from pytorch_connectomics_b200.architectures import monai_unet as MU
    from pytorch_connectomics_b200.training.optim import reference_param_groups
    B = ref_loader.ref_optimizer_build()
    kw = dict(exp_r=2, kernel_size=3, deep_supervision=True, do_res=True, do_res_up_down=True, block_counts=[1] * 9)
    shared = torch.nn.Sequential(torch.nn.Linear(3, 3), torch.nn.Linear(3, 3))
    shared[1].weight = shared[0].weight
    nets = [PM.MedNeXt(1, 16, 2, **kw), PM.MedNeXt(1, 16, 2, norm_type="layer", **kw),
            MU.UNet(3, 1, 2, [16, 32, 64], [2, 2], num_res_units=1), shared,
            A.build_model(NS(model=NS(arch=NS(type="mednext"), in_channels=1, out_channels=2, heads=None,
                                      mednext=NS(size="S", kernel_size=3), loss=NS(deep_supervision=False))))]
    nets[0].stem.bias.requires_grad_(False)                         # frozen parameters are left out by both
    for oc in (NS(name="adamw", lr=1e-3, weight_decay=0.01),
               NS(name="AdamW", lr=3e-4, weight_decay=0.05, weight_decay_norm=0.001, weight_decay_bias=0.0, bias_lr_factor=2.0)):
        cfg = NS(optimization=NS(optimizer=oc))
        for net in nets:
            real = B.build_optimizer(cfg, net).param_groups
            ours = reference_param_groups(net, oc.lr, oc.weight_decay, getattr(oc, "weight_decay_norm", 0.0),
                                          getattr(oc, "weight_decay_bias", None), getattr(oc, "bias_lr_factor", 1.0))
            assert len(real) == len(ours) > 0
            for a, b in zip(real, ours):
                assert a["params"][0] is b["params"][0] and a["lr"] == b["lr"] and a["weight_decay"] == b["weight_decay"]
    layer_net = nets[1]
    g = {id(x["params"][0]): x for x in reference_param_groups(layer_net, 1e-3, 0.01, 0.0, 0.002, 1.0)}
    blk = layer_net.enc_block_0[0]
    assert g[id(blk.norm.weight)]["weight_decay"] == 0.01 and g[id(blk.norm.bias)]["weight_decay"] == 0.002


def test_artifact_metadata_equals_the_real_artifact_module():
    """`build_prediction_artifact_metadata` and the attrs encoding (`inference/artifact.py:75-138`) against the REAL module
    executed in place: same dataclass fields and the same JSON / scalar attrs on the dataset."""
    from dataclasses import asdict
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference is only present in the build container")
    from pytorch_connectomics_b200.inference import artifact as PA
    RA = ref_loader.ref_artifact()
    cfgs = [NS(),
            NS(model=NS(arch=NS(type="mednext"), primary_head="aff"), data=NS(data_transform=NS(val_transpose=[2, 1, 0])),
               decoding=NS(enabled=False), inference=NS(model=NS(select_channel="0:3"),
                                                       prediction_transform=NS(enabled=True, intensity_scale=255, intensity_dtype="uint8"))),
            NS(model=NS(arch=NS(type="monai_unet"), primary_head=None), data=NS(data_transform=NS(val_transpose=[])),
               inference=NS(model=NS(select_channel=[2, 0]), prediction_transform=NS(enabled=False, intensity_scale=3.0)))]
    calls = [dict(),
             dict(image_path="/d/img.h5", checkpoint_path="/c/last.ckpt", output_head="sdt", input_shape=[12, 10, 14], final_shape=(10, 10, 10),
                  crop_pad=((1, 1), (0, 0), [2, 2]), chunk_shape=[6, 10, 7], halo=(2, 2, 2), extra={"compression": "gzip", "chunk_index_zyx": [0, 0, 1]}),
             dict(input_shape=[], intensity_scale=0.5, intensity_dtype="float16", extra=None)]

    class Dset:
        def __init__(self):
            self.attrs = {}

    for cfg in cfgs:
        for kw in calls:
            want, got = RA.build_prediction_artifact_metadata(cfg, **kw), PA.build_prediction_artifact_metadata(cfg, **kw)
            assert asdict(want) == asdict(got), (kw, asdict(want), asdict(got))
            d = Dset()
            RA.write_prediction_artifact_attrs(d, want)
            assert d.attrs == PA.metadata_attrs(got)

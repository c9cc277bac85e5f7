"""z-slab sharded sliding-window inference (pytorch_connectomics_b200/inference/sharded.py).

CPU: the integer plan (every eager-grid window exactly once, own ranges partition the volume, sends mirror
recvs; the C5 2048^3/160/0.5 geometry) and the overlap exchange over gloo at world size 2, with the oracle's
accumulate emulating the per-rank kernel phase.  GPU: the real per-rank phase for three simulated ranks on one
device against EagerSlidingWindowEngine (bit-exact outside the exchanged planes, 1-ulp-level inside)."""
import os

import pytest
import torch

from oracle import window_oracle as O
from pytorch_connectomics_b200.inference import sharded as S


def _check_plans(plans, image, roi, overlap):
    starts = O.dense_starts(image, roi, O.scan_interval(image, roi, overlap))
    got = [w for p in plans for w in p.windows]
    assert sorted(got) == sorted(starts) and len(got) == len(set(got))        # every window exactly once
    assert [w for p in plans for w in p.windows] == starts                    # and in grid order rank by rank
    live = [p for p in plans if p.windows]
    assert live[0].own[0] == 0 and live[-1].own[1] == image[0]
    for a, b in zip(live, live[1:]):
        assert a.own[1] == b.own[0]                                           # own ranges partition [0, D)
    for p in live:
        assert p.slab[0] <= p.own[0] < p.own[1] <= p.slab[1]
        assert p.slab == (p.z_starts[0], p.z_starts[-1] + roi[0])
    sends = sorted((p.rank, q, lo, hi) for p in plans for q, lo, hi in p.sends)
    recvs = sorted((q, p.rank, lo, hi) for p in plans for q, lo, hi in p.recvs)
    assert sends == recvs
    # every plane a rank accumulates but does not own is sent to its owner, nothing else is
    for p in live:
        foreign = set(range(p.slab[0], p.slab[1])) - set(range(p.own[0], p.own[1]))
        sent = set()
        for q, lo, hi in p.sends:
            assert plans[q].own[0] <= lo < hi <= plans[q].own[1]
            sent |= set(range(lo, hi))
        assert sent == foreign


@pytest.mark.parametrize("image,roi,overlap,world", [
    ((48, 40, 40), (16, 16, 16), 0.5, 2), ((48, 40, 40), (16, 16, 16), 0.5, 3), ((50, 33, 20), (16, 16, 16), 0.25, 4),
    ((40, 16, 16), (16, 16, 16), 0.75, 3), ((16, 16, 16), (16, 16, 16), 0.5, 2), ((33, 20, 20), (16, 8, 8), 0.5, 8),
])
def test_plan_z_slabs_partitions_the_eager_grid(image, roi, overlap, world):
    plans = S.plan_z_slabs(image, roi, overlap, world)
    assert len(plans) == world
    _check_plans(plans, image, roi, overlap)


def test_plan_c5_geometry():
    # BASELINE configs[4]: 2048^3, 160^3 tiles, 50 % overlap, 8 GPUs -> 25 z-starts split 4,3,3,3,3,3,3,3
    plans = S.plan_z_slabs((2048,) * 3, (160,) * 3, 0.5, 8)
    assert [len(p.z_starts) for p in plans] == [4, 3, 3, 3, 3, 3, 3, 3]
    assert sum(len(p.windows) for p in plans) == 15625
    assert plans[0].slab == (0, 400) and plans[0].own == (0, 320)
    assert plans[7].z_starts == [1760, 1840, 1888] and plans[7].own == (1760, 2048)
    for p in plans[:-1]:
        assert p.sends == [(p.rank + 1, p.slab[1] - 80, p.slab[1])] or p.rank == 6   # 80 overlap planes per face
    assert S.split_contiguous(25, 8) == [(0, 4), (4, 7), (7, 10), (10, 13), (13, 16), (16, 19), (19, 22), (22, 25)]
    with pytest.raises(ValueError):
        S.plan_z_slabs((100, 100, 100), (160, 160, 160), 0.5, 2)


def _cpu_rank_phase(vol, net, plan, mode):
    """oracle emulation of ZSlabShardedEngine.accumulate_local (test-only; the product path is CUDA)."""
    roi = plan.roi
    z0, z1 = plan.slab
    w = O.importance_map(roi, mode, dtype=torch.float32).view(1, 1, *roi)
    val = torch.zeros((1, 2, z1 - z0, *plan.image[1:]))
    wacc = torch.zeros((1, 1, z1 - z0, *plan.image[1:]))
    for s in plan.windows:
        out = net(O.extract_patch(vol, s, roi, "constant", 0.0))
        idx = (slice(None), slice(None), slice(s[0] - z0, s[0] - z0 + roi[0]), slice(s[1], s[1] + roi[1]),
               slice(s[2], s[2] + roi[2]))
        val[idx] += out * w
        wacc[idx] += w
    return val, wacc


def _net(x):   # deterministic 1 -> 2 channel "network"
    return torch.cat([x * 0.5 + 0.25, torch.sin(x * 3.0)], dim=1)


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(5)
    vol = torch.rand(1, 1, 44, 24, 20)
    plan = S.plan_z_slabs(vol.shape[2:], (16, 16, 16), 0.5, world)[rank]
    val, wacc = _cpu_rank_phase(vol, _net, plan, "bump")
    # the comm= route (what NativeComm / pcb_sw_exchange_overlap serves on the GPU) through a stand-in with the same
    # exchange(sends, recvs) contract carried by gloo: same packed messages, same result as the torch.distributed route
    val2, wacc2 = val.clone(), wacc.clone()

    class GlooComm:
        rank, world = dist.get_rank(), dist.get_world_size()

        @staticmethod
        def exchange(sends, recvs):
            ops = [dist.P2POp(dist.isend, t, peer) for t, peer in sends] + [dist.P2POp(dist.irecv, t, peer) for t, peer in recvs]
            for req in (dist.batch_isend_irecv(ops) if ops else []):
                req.wait()

    S.exchange_overlaps(val2, wacc2, plan, comm=GlooComm)
    S.exchange_overlaps(val, wacc, plan)
    assert torch.equal(val, val2) and torch.equal(wacc, wacc2)
    z0 = plan.slab[0]
    own = O.normalize_accumulator(val[:, :, plan.own[0] - z0:plan.own[1] - z0].clone(),
                                  wacc[:, :, plan.own[0] - z0:plan.own[1] - z0].clone())
    q.put((rank, plan.own, own.numpy()))     # by value: torch tensors travel as fds that die with the worker
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_overlaps_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    from conftest import free_port
    port = free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    torch.manual_seed(5)
    vol = torch.rand(1, 1, 44, 24, 20)
    want = O.eager_sliding_window(vol, _net, (16, 16, 16), overlap=0.5, mode="bump")
    got = torch.cat([torch.from_numpy(r[2]) for r in res], dim=2)
    assert got.shape == want.shape
    assert [r[1] for r in res][0][0] == 0 and res[-1][1][1] == 44
    assert torch.allclose(got, want, rtol=2e-6, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("world,overlap", [(1, 0.5), (3, 0.5), (2, 0.25)])
def test_sharded_engine_matches_eager_gpu(world, overlap):
    from pytorch_connectomics_b200.inference.window import EagerSlidingWindowEngine
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    vol = torch.rand(1, 1, 52, 40, 36, device=dev)
    kw = dict(roi_size=(16, 16, 16), sw_batch_size=3, overlap=overlap, mode="bump", padding_mode="constant", cval=0.0)
    want = EagerSlidingWindowEngine(**kw)(inputs=vol, network=_net)
    plans = S.plan_z_slabs(vol.shape[2:], (16, 16, 16), overlap, world)
    engines = [S.ZSlabShardedEngine(device=dev, rank=r, world=world, **kw) for r in range(world)]
    if world == 1:
        got, own = engines[0](vol, _net)
        assert own == (0, 52) and torch.equal(got, want)          # same order of sums -> bit-identical
        return
    acc = [e.accumulate_local(vol.cpu(), _net, p) for e, p in zip(engines, plans)]   # host volume: slab staged per rank
    for p in plans:                                               # in-process stand-in for the NCCL send/recv pairs
        for peer, lo, hi in p.recvs:
            src_v, src_w = acc[peer]
            z0s, z0d = plans[peer].slab[0], p.slab[0]
            acc[p.rank][0][0, :, lo - z0d:hi - z0d] += src_v[0, :, lo - z0s:hi - z0s]
            acc[p.rank][1][0, :, lo - z0d:hi - z0d] += src_w[0, :, lo - z0s:hi - z0s]
    parts = [S.ZSlabShardedEngine.finalize(v, w, p) for (v, w), p in zip(acc, plans)]
    got = torch.cat(parts, dim=2)
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=2e-6, atol=1e-6)
    exchanged = torch.zeros(52, dtype=torch.bool)
    for p in plans:
        for _, lo, hi in p.recvs:
            exchanged[lo:hi] = True
    keep = (~exchanged).nonzero().flatten().to(dev)
    assert torch.equal(got.index_select(2, keep), want.index_select(2, keep))   # untouched planes are bit-exact


def _nccl_worker(rank, world, port, q):
    """one process per GPU over NCCL: the product path of ZSlabShardedEngine (send/recv of the overlap planes)"""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(7)
    vol = torch.rand(1, 1, 52, 40, 36)
    kw = dict(roi_size=(16, 16, 16), sw_batch_size=3, overlap=0.5, mode="bump", padding_mode="constant", cval=0.0)
    eng = S.ZSlabShardedEngine(device=dev, **kw)
    part, own = eng(vol, _net)                                  # host volume: only this rank's slab goes to its GPU
    plan = S.plan_z_slabs(vol.shape[2:], (16, 16, 16), 0.5, world)[rank]
    part2 = eng.run_slab(vol[:, :, plan.slab[0]:plan.slab[1]].to(dev), _net, plan)     # pre-sliced device slab
    assert torch.equal(part, part2)
    q.put((rank, own, part.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_engine_two_ranks_nccl():
    """2 GPUs, 2 processes, NCCL: the exchanged result equals the single-GPU eager engine (bit-exact away from the
    exchanged planes).  Skipped on a single-GPU box (run with `gpurun --gpus 2`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pytorch_connectomics_b200.inference.window import EagerSlidingWindowEngine
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    from conftest import free_port
    port = free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
    torch.manual_seed(7)
    vol = torch.rand(1, 1, 52, 40, 36)
    kw = dict(roi_size=(16, 16, 16), sw_batch_size=3, overlap=0.5, mode="bump", padding_mode="constant", cval=0.0)
    want = EagerSlidingWindowEngine(**kw)(inputs=vol.to("cuda:0"), network=_net).cpu()
    got = torch.cat([torch.from_numpy(r[2]) for r in res], dim=2)
    assert got.shape == want.shape and res[0][1][0] == 0 and res[1][1][1] == 52
    assert torch.allclose(got, want, rtol=2e-6, atol=1e-6)

"""The DDP seam as an object (``training/data_parallel.py``; reference ``training/lightning/trainer.py:231-256,314-334``):
``ArenaDataParallel`` must give what Lightning's ``DDPStrategy(find_unused_parameters=True)`` + ``accumulate_grad_batches`` give —
mean gradients after ``backward()`` alone, tolerance of unused parameters (also when they differ between ranks), ``no_sync``
accumulation, ``register_comm_hook`` with DDP's hook protocol — with the exchange launched per arena segment while the
backward pass still runs.  Host logic only: world-size-2 gloo on the CPU; the in-process single-rank results are the oracle."""

import os

import pytest
import torch

from pytorch_connectomics_b200.training.data_parallel import plan_segments


class _Net(torch.nn.Sequential):
    """layer 4 is never used (MedNeXt's unused deep-supervision heads); layer 2 can be skipped (on rank 1 only in the test)"""

    def __init__(self):
        super().__init__(torch.nn.Linear(4, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Linear(16, 3),
                         torch.nn.Linear(3, 3))

    def forward(self, x, skip_mid=False, everything=False):
        if everything:
            return super().forward(x)
        h = self[1](self[0](x))
        if not skip_mid:
            h = h + self[2](h)
        return {"output": self[3](h)}


def _net():
    return _Net()


def _forward(net, x, skip_mid):
    return net(x, skip_mid)["output"]


def _data(rank, step):
    g = torch.Generator().manual_seed(100 * step + rank)
    return torch.randn(5, 4, generator=g)


def _flat_grads(net):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in net.parameters()])


def test_plan_segments_properties():
    numels = [10, 3, 64, 1, 1, 200, 7]
    align = 8
    offsets, off = [], 0
    for n in numels:
        offsets.append(off)
        off += -(-n // align) * align
    plan = plan_segments(numels, offsets, off, cap_elems=60, first_cap_elems=5)
    # launch order walks from the last parameter to the first; the groups tile the parameter list and the arena exactly
    assert plan[0][1] == len(numels) and plan[-1][0] == 0
    assert all(a[0] == b[1] for a, b in zip(plan, plan[1:]))
    assert all(a[2] == b[3] for a, b in zip(plan, plan[1:])) and plan[0][3] == off and plan[-1][2] == 0
    assert plan[0][:2] == (6, 7)                                         # small first bucket: just the last parameter
    for first, last, lo, hi in plan:
        assert lo == offsets[first] and hi == (offsets[last] if last < len(numels) else off)
    assert sum(sum(numels[a:b]) for a, b, _, _ in plan) == sum(numels)
    one = plan_segments(numels, offsets, off, cap_elems=10 ** 9)
    assert one == [(0, len(numels), 0, off)]
    assert plan_segments([], [], 0, 4) == []


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception as e:                      # fail the parent quickly instead of letting it wait for the queue
        import traceback
        q.put((rank, {"error": f"{e!r}\n{traceback.format_exc()}"}))


def _worker_body(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorch_connectomics_b200.training import (ArenaDataParallel, allreduce_sum_hook, bf16_compress_hook)
    out = {}
    torch.manual_seed(7 + rank)                                  # different initial weights: rank 0's must win
    net = _net()
    kb = 1.0 / 1024
    adp = ArenaDataParallel(net, bucket_cap_mb=0.5 * kb, first_bucket_mb=0.01 * kb)
    out["params"] = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy()
    out["n_segments"] = len(adp.segments)

    # 1. backward() alone leaves the mean; Lightning's zero_grad(set_to_none=True) detaches the views on rank 1 first
    if rank == 1:
        net.zero_grad(set_to_none=True)
    else:
        adp.zero_grad()
    _forward(adp, _data(rank, 0), skip_mid=rank == 1).square().sum().backward()
    out["mean"] = _flat_grads(net).numpy()
    out["log"] = list(adp.launch_log)
    out["views"] = all(p.grad.data_ptr() == adp.arena.view_of(i).data_ptr() for i, p in enumerate(adp.arena.params))

    # 2. accumulate_grad_batches = 2: nothing is exchanged inside no_sync(), the boundary backward exchanges the window's sum
    adp.zero_grad()
    with adp.no_sync():
        _forward(adp, _data(rank, 1), False).square().sum().backward()
    out["local_after_no_sync"] = _flat_grads(net).numpy()
    _forward(adp, _data(rank, 2), False).square().sum().backward()
    out["accum"] = _flat_grads(net).numpy()

    # 3. reduce_op='sum' semantics through the hook API, on a second wrapper over the SAME arena
    adp.remove_hooks()
    adp2 = ArenaDataParallel(net, arena=adp.arena, reduce_op="sum", bucket_cap_mb=0.5 * kb, init_sync=False)
    adp2.zero_grad()
    _forward(adp2, _data(rank, 3), False).square().sum().backward()
    out["sum"] = _flat_grads(net).numpy()

    # 4. a compression hook (returns through a different tensor), and the record of the buckets it saw
    seen = []

    def hook(state, bucket):
        seen.append((bucket.index(), bucket.is_last(), bucket.buffer().numel(), len(bucket.parameters()), len(bucket.gradients())))
        return bf16_compress_hook(state, bucket)

    adp2.register_comm_hook(None, hook)
    adp2.zero_grad()
    _forward(adp2, _data(rank, 4), False).square().sum().backward()
    out["bf16"] = _flat_grads(net).numpy()
    out["seen"] = seen

    # 5. reduce_now(): gradients that were produced without hooks (a CUDA-graph replay) are exchanged on request
    adp2.remove_hooks()
    adp3 = ArenaDataParallel(net, arena=adp.arena, overlap=False, init_sync=False)
    adp3.zero_grad()
    with adp3.no_sync():
        _forward(adp3, _data(rank, 5), False).square().sum().backward()
    adp3.reduce_now()
    out["reduce_now"] = _flat_grads(net).numpy()
    adp3.remove_hooks()

    # 5b. stale arena: after a full step, Lightning's zero_grad(set_to_none=True) detaches every view WITHOUT clearing the
    # arena; a parameter that gets no gradient in the next step (layer 2 on rank 1) must contribute zero, not the old slice,
    # and the repair must not touch a segment that has already been exchanged
    adp4 = ArenaDataParallel(net, arena=adp.arena, bucket_cap_mb=0.5 * kb, first_bucket_mb=0.01 * kb, init_sync=False)
    adp4.zero_grad()
    _forward(adp4, _data(rank, 7), False).square().sum().backward()
    net.zero_grad(set_to_none=True)
    _forward(adp4, _data(rank, 8), skip_mid=rank == 1).square().sum().backward()
    out["stale"] = _flat_grads(net).numpy()
    out["stale_views"] = all(p.grad is not None and p.grad.data_ptr() == adp.arena.view_of(i).data_ptr()
                             for i, p in enumerate(adp.arena.params))
    adp4.remove_hooks()

    # 6. the same sum hook under torch's own DistributedDataParallel (GradBucket protocol)
    torch.manual_seed(3)
    ddp_net = _net()
    ddp = torch.nn.parallel.DistributedDataParallel(ddp_net, find_unused_parameters=True)
    ddp.register_comm_hook(None, allreduce_sum_hook)
    y = ddp(_data(rank, 6), everything=True)                                     # DDP needs its own forward to arm the reducer
    y.square().sum().backward()
    out["ddp_sum"] = _flat_grads(ddp_net).numpy()
    q.put((rank, out))
    dist.destroy_process_group()


def _local(step_ranks, skip=lambda r: False, seed=7, fwd=_forward):
    """per-rank gradients of rank 0's weights, computed in-process: [(flat grad of rank r)]"""
    torch.manual_seed(seed)
    net = _net()
    gs = []
    for r, steps in step_ranks:
        net.zero_grad()
        for s in steps:
            fwd(net, _data(r, s), skip(r)).square().sum().backward()
        gs.append(_flat_grads(net).clone())
    return net, gs


@pytest.mark.timeout(600)
def test_arena_data_parallel_gloo_world2():
    import torch.multiprocessing as mp
    from conftest import free_port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, o = q.get(timeout=300)
        if "error" in o:
            for p in procs:
                p.kill()
            pytest.fail(f"rank {r}: {o['error']}")
        got[r] = o
    for p in procs:
        p.join(timeout=60)
    t = lambda r, k: torch.from_numpy(got[r][k])
    # construction: rank 0's weights everywhere; several segments so the overlap logic is exercised
    net, gs = _local([(0, [0]), (1, [0])], skip=lambda r: r == 1)
    ref_params = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    assert torch.equal(t(0, "params"), ref_params) and torch.equal(t(1, "params"), ref_params)
    assert got[0]["n_segments"] >= 3 and got[0]["n_segments"] == got[1]["n_segments"]
    # 1. mean gradient, identical on both ranks, unused parameters zero, views re-attached
    assert torch.equal(t(0, "mean"), t(1, "mean"))
    assert torch.allclose(t(0, "mean"), (gs[0] + gs[1]) / 2, rtol=1e-6, atol=1e-7)
    assert t(0, "mean")[-12:].abs().sum() == 0 and got[0]["views"] and got[1]["views"]
    n_params = len(list(net.parameters()))
    for r in range(2):
        log = got[r]["log"]
        assert [k for k, _ in log] == list(range(got[r]["n_segments"]))       # same launch order on every rank
    # overlap: forward() marked the dead head (and, on rank 1, the skipped layer) ready, so the first segments go out while the
    # backward pass has produced only the gradients of the last used layer
    assert got[0]["log"][0][1] <= 4 and got[0]["log"][1][1] < n_params and got[1]["log"][0][1] <= 6
    # 2. no_sync leaves local gradients; the boundary step exchanges the sum of the window
    _, g1 = _local([(0, [1]), (1, [1])])
    assert torch.allclose(t(0, "local_after_no_sync"), g1[0], rtol=1e-6, atol=1e-7)
    assert torch.allclose(t(1, "local_after_no_sync"), g1[1], rtol=1e-6, atol=1e-7)
    _, g12 = _local([(0, [1, 2]), (1, [1, 2])])
    assert torch.equal(t(0, "accum"), t(1, "accum"))
    assert torch.allclose(t(0, "accum"), (g12[0] + g12[1]) / 2, rtol=1e-5, atol=1e-6)
    # 3. SUM for the fused optimizer
    _, g3 = _local([(0, [3]), (1, [3])])
    assert torch.allclose(t(0, "sum"), g3[0] + g3[1], rtol=1e-6, atol=1e-6) and torch.equal(t(0, "sum"), t(1, "sum"))
    # 4. bf16 compression hook: mean to bf16 accuracy; the hook saw every segment once, in order, last one flagged
    _, g4 = _local([(0, [4]), (1, [4])])
    want = (g4[0] + g4[1]) / 2
    assert (t(0, "bf16") - want).norm() <= 1e-2 * want.norm() and torch.equal(t(0, "bf16"), t(1, "bf16"))
    seen = got[0]["seen"]
    assert [s[0] for s in seen] == list(range(len(seen))) and [s[1] for s in seen] == [False] * (len(seen) - 1) + [True]
    assert sum(s[3] for s in seen) == n_params and all(s[3] == s[4] for s in seen)
    # 5. reduce_now
    _, g5 = _local([(0, [5]), (1, [5])])
    assert torch.allclose(t(0, "reduce_now"), (g5[0] + g5[1]) / 2, rtol=1e-6, atol=1e-7)
    # 5b. stale arena + detached views
    _, g8 = _local([(0, [8]), (1, [8])], skip=lambda r: r == 1)
    assert torch.equal(t(0, "stale"), t(1, "stale")) and got[0]["stale_views"] and got[1]["stale_views"]
    assert torch.allclose(t(0, "stale"), (g8[0] + g8[1]) / 2, rtol=1e-6, atol=1e-7)
    # 6. torch DDP + our sum hook
    _, g6 = _local([(0, [6]), (1, [6])], seed=3, fwd=lambda net, x, _s: net(x, everything=True))
    assert torch.allclose(t(0, "ddp_sum"), g6[0] + g6[1], rtol=1e-5, atol=1e-6)


def test_single_process_wrapper_needs_no_process_group():
    from pytorch_connectomics_b200.training import ArenaDataParallel
    torch.manual_seed(0)
    net = _net()
    adp = ArenaDataParallel(net, bucket_cap_mb=1e-4)
    _forward(adp, _data(0, 0), False).square().sum().backward()
    got = _flat_grads(net).clone()
    _, gs = _local([(0, [0])], seed=0)
    assert torch.equal(got, gs[0]) and len(adp.launch_log) == len(adp.segments)
    with pytest.raises(TypeError):
        adp.register_comm_hook(None, "not callable")
    with pytest.raises(ValueError):
        ArenaDataParallel(_net(), reduce_op="max")
    with pytest.raises(ValueError):
        ArenaDataParallel(_net(), arena=adp.arena)


def test_static_graph_caches_the_unused_set_and_detects_a_changed_graph():
    from pytorch_connectomics_b200.training import ArenaDataParallel
    torch.manual_seed(0)
    adp = ArenaDataParallel(_net(), bucket_cap_mb=1e-6, first_bucket_mb=None, static_graph=True)
    assert len(adp.segments) == 8                                # one parameter per segment at this cap
    _forward(adp, _data(0, 0), True).square().sum().backward()   # the first iteration skips layer 2 ...
    assert adp._static_unused == [2, 3, 6, 7]
    adp.zero_grad()
    with pytest.raises(RuntimeError, match="after its segment was exchanged"):
        _forward(adp, _data(0, 1), False).square().sum().backward()   # ... the second one uses it: loud, not silently stale


def test_lightning_strategy_factory_against_a_stub_ddp_strategy(monkeypatch):
    """Lightning is not in this image: the factory is exercised against a stub of the three ``DDPStrategy`` members it
    overrides (``_setup_model``, ``_register_ddp_hooks``, ``block_backward_sync``) so that at least the class body, the
    keyword routing and the no_sync plumbing run.  Without any Lightning the factory raises ImportError."""
    import sys
    import types
    from pytorch_connectomics_b200.training import ArenaDataParallel, allreduce_sum_hook, make_arena_ddp_strategy
    for name in ("lightning", "lightning.pytorch", "lightning.pytorch.strategies", "pytorch_lightning", "pytorch_lightning.strategies"):
        monkeypatch.setitem(sys.modules, name, None)
    with pytest.raises(ImportError):
        make_arena_ddp_strategy()

    class DDPStrategy:
        def __init__(self, **kw):
            self._ddp_kwargs = {"bucket_cap_mb": 0.001, "find_unused_parameters": True, "unknown": 1}
            self.init_kw = kw
            self.model = None

    pl = types.ModuleType("pytorch_lightning")
    st = types.ModuleType("pytorch_lightning.strategies")
    st.DDPStrategy = DDPStrategy
    pl.strategies = st
    monkeypatch.setitem(sys.modules, "pytorch_lightning", pl)
    monkeypatch.setitem(sys.modules, "pytorch_lightning.strategies", st)
    strat = make_arena_ddp_strategy({"timeout": 5}, reduce_op="sum", init_sync=False)
    assert isinstance(strat, DDPStrategy) and strat.init_kw == {"timeout": 5}
    strat.model = strat._setup_model(_net())
    assert isinstance(strat.model, ArenaDataParallel) and strat.model.reduce_op == "sum" and len(strat.model.segments) > 1
    strat._ddp_comm_hook, strat._ddp_comm_state = allreduce_sum_hook, None
    strat._register_ddp_hooks()
    assert strat.model._hook is allreduce_sum_hook
    with strat.block_backward_sync():
        assert strat.model.require_backward_grad_sync is False
    assert strat.model.require_backward_grad_sync is True


def test_arena_train_step_drives_the_wrapper(monkeypatch):
    """``ArenaTrainStep`` over ``ArenaDataParallel`` (host logic; the fused optimizer kernel is replaced by a recorder): the
    micro-batches before the boundary run under ``no_sync()``, the boundary backward exchanges, ``flush()`` exchanges through
    ``reduce_now()``, and the optimizer is told whether the gradients are sums (``reduce_op='sum'``) or already means."""
    from pytorch_connectomics_b200.training import ArenaDataParallel, ArenaTrainStep, FlatGradArena, FusedAdamW
    for reduce_op in ("sum", "mean"):
        torch.manual_seed(0)
        net = _net()
        arena = FlatGradArena(net.parameters())
        opt = FusedAdamW.__new__(FusedAdamW)             # no CUDA here: only the attributes the step touches
        opt.arena = arena
        steps = []
        opt.step = lambda grads_are_summed=False, _s=steps, _a=arena: _s.append((grads_are_summed, _a.buffer.clone()))
        ddp = ArenaDataParallel(net, arena=arena, reduce_op=reduce_op, bucket_cap_mb=1e-4, first_bucket_mb=None)
        synced = []
        orig = ddp._finalize
        monkeypatch.setattr(ddp, "_finalize", lambda: (synced.append(len(steps)), orig())[1])
        step = ArenaTrainStep(ddp, lambda out, t: (out["output"] - t).square().mean(), opt, accumulate_grad_batches=2)
        xs = [_data(0, s) for s in range(3)]
        t = torch.zeros(5, 3)
        step(xs[0], t)
        assert steps == [] and synced == []              # first micro-batch: no exchange, no optimizer step
        step(xs[1], t)
        assert len(steps) == 1 and synced == [0] and steps[0][0] == (reduce_op == "sum")
        # the window's gradient is the sum of the two micro-batch gradients of loss / 2
        ref = _net()
        ref.load_state_dict(net.state_dict())
        for x in xs[:2]:
            ((ref(x)["output"] - t).square().mean() / 2).backward()
        assert torch.allclose(steps[0][1][:arena.total], torch.cat([
            torch.nn.functional.pad((p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1),
                                    (0, (arena.offsets[i + 1] if i + 1 < len(arena.params) else arena.total) - arena.offsets[i] - p.numel()))
            for i, p in enumerate(ref.parameters())]), atol=1e-6)
        step(xs[2], t)                                   # a partial window ...
        assert len(steps) == 1 and step.flush() == 1     # ... is stepped by flush(), through reduce_now()
        assert len(steps) == 2 and synced == [0, 1] and step.flush() is None
    with pytest.raises(ValueError, match="optimizer's gradient arena"):
        ArenaTrainStep(ArenaDataParallel(_net()), lambda o, t: o, opt)

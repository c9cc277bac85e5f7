"""Discrete-event model of the mbarrier protocols of `mlp_bwd_ws_kernel` / `mlp_bwd_ws2_kernel` (csrc/mednext_bwd.cu) and
`mlp_fused_kernel` (csrc/mednext_fwd.cu).

The warp-specialised backward kernels hand tiles between four loader warps, one MMA-issuing thread and two epilogue
groups through mbarriers only.  A wrong phase parity or barrier index does not fail loudly on the GPU — it races or hangs —
so the protocol is restated here, role by role with the SAME index / parity formulas as the CUDA code, and run under
randomised timings with a data-hazard checker:

* every shared resource (operand stage sA/sD, accumulators acc1/accG/accD, sH/sDh) records which tile it holds and who is
  still reading it; a write while readers are outstanding, or a read of the wrong tile, raises;
* `mbarrier.try_wait.parity` has the hardware semantics (succeeds iff the barrier's current phase parity differs from the
  requested one), so a waiter that skipped a completion — parity aliasing — is caught as a wrong-tile read or a deadlock;
* tcgen05 MMAs execute asynchronously and in order; `tcgen05.commit` arrives when everything issued before it has retired.

    python tools/ws_protocol_model.py            # all configurations x 200 random schedules

`tests/test_ws_protocol.py` runs a reduced sweep on CPU."""
from __future__ import annotations

import heapq
import random
import sys
from typing import Dict, List, Optional


class Hazard(AssertionError):
    pass


class Barrier:
    def __init__(self, name: str, count: int):
        self.name, self.count, self.pending, self.phase = name, count, 0, 0

    def arrive(self):
        self.pending += 1
        if self.pending > self.count:
            raise Hazard(f"{self.name}: more arrivals than the expected count")
        if self.pending == self.count:
            self.pending, self.phase = 0, self.phase + 1

    def test(self, parity: int) -> bool:          # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


class Resource:
    """One buffer: which tile it holds, whether the write has landed, and the reads still outstanding."""

    def __init__(self, name: str):
        self.name, self.tile, self.valid, self.readers, self.writing = name, None, False, 0, False

    def begin_write(self, tile: int):
        if self.readers:
            raise Hazard(f"{self.name}: write of tile {tile} while {self.readers} read(s) of tile {self.tile} are outstanding")
        if self.writing:
            raise Hazard(f"{self.name}: two writers (tile {tile})")
        self.writing, self.valid, self.tile = True, False, tile

    def end_write(self, tile: int):
        assert self.writing and self.tile == tile
        self.writing, self.valid = False, True

    def begin_read(self, tile: int, who: str):
        if not self.valid or self.tile != tile:
            raise Hazard(f"{self.name}: {who} expects tile {tile}, buffer holds {self.tile} (valid={self.valid})")
        self.readers += 1

    def end_read(self, tile: int):
        assert self.readers > 0
        self.readers -= 1


class Sim:
    def __init__(self, ntiles: int, NB: int, NST: int, seed: int):
        self.ntiles, self.NB, self.NST = ntiles, NB, NST
        self.rng = random.Random(seed)
        self.now = 0.0
        self.events: List = []
        self.seq = 0
        self.bars: Dict[str, List[Barrier]] = {
            "a_full": [Barrier(f"a_full[{i}]", 1) for i in range(4)],     # 32 lanes of ONE warp arrive: modelled as 1
            "a_empty": [Barrier(f"a_empty[{i}]", 1) for i in range(4)],
            "hp_full": [Barrier(f"hp_full[{i}]", 1) for i in range(2)],
            "e1_done": [Barrier(f"e1_done[{i}]", 1) for i in range(2)],   # 128 threads of ONE group: modelled as 1
            "d_full": [Barrier(f"d_full[{i}]", 1) for i in range(2)],
            "d_empty": [Barrier(f"d_empty[{i}]", 1) for i in range(2)],
            "h_free": [Barrier(f"h_free[{i}]", 1) for i in range(2)],
        }
        self.sAD = [Resource(f"sA/sD[{i}]") for i in range(NST)]
        self.acc = [Resource(f"acc1/accG[{i}]") for i in range(NB)]
        self.sH = [Resource(f"sH/sDh[{i}]") for i in range(NB)]
        self.accD = [Resource(f"accD[{i}]") for i in range(NB)]
        self.waiting: List = []               # (generator, barrier, parity)
        self.mma_free_at = 0.0                # the tensor pipe executes in order
        self.done_tiles_e2 = 0
        self.roles_alive = 0

    # ---- scheduling helpers
    def at(self, t: float, fn):
        self.seq += 1
        heapq.heappush(self.events, (t, self.seq, fn))

    def dur(self, lo: float, hi: float) -> float:
        return self.rng.uniform(lo, hi)

    def run_role(self, gen):
        """Advance a role generator until it blocks; it yields ('wait', bar, parity) or ('sleep', dt)."""
        try:
            while True:
                op = next(gen)
                if op[0] == "sleep":
                    self.at(self.now + op[1], lambda g=gen: self.run_role(g))
                    return
                if op[0] == "wait":
                    _, bar, parity = op
                    if bar.test(parity):
                        continue
                    self.waiting.append((gen, bar, parity))
                    return
                raise ValueError(op)
        except StopIteration:
            self.roles_alive -= 1

    def wake(self):
        still = []
        ready = []
        for gen, bar, parity in self.waiting:
            (ready if bar.test(parity) else still).append((gen, bar, parity))
        self.waiting = still
        for gen, _, _ in ready:
            self.run_role(gen)

    def arrive(self, bar: Barrier):
        bar.arrive()
        self.wake()

    def mma(self, duration: float, reads, writes, tile_r: Dict[Resource, int], commits: List[Barrier]):
        """Issue an asynchronous MMA group now: executes after everything issued before, reads/writes checked at start/end."""
        start = max(self.now, self.mma_free_at)
        end = start + duration
        self.mma_free_at = end

        def begin():
            for r in reads:
                r.begin_read(tile_r[r], "MMA")
            for w in writes:
                w.begin_write(tile_r[w])

        def finish():
            for r in reads:
                r.end_read(tile_r[r])
            for w in writes:
                w.end_write(tile_r[w])
            for b in commits:
                self.arrive(b)

        self.at(start, begin)
        self.at(end, finish)

    # ---- roles (formulas copied from mlp_bwd_ws2_kernel; NB = 2, NST = 4 is mlp_bwd_ws_kernel)
    def loader(self, warp: int):
        NST, B = self.NST, self.bars
        it = warp
        while it < self.ntiles:
            s = it % NST
            prev = it - NST
            if prev >= 0:
                yield ("wait", B["a_empty"][prev & 3], (prev >> 2) & 1)
            self.sAD[s].begin_write(it)
            yield ("sleep", self.dur(0.5, 3.0))
            self.sAD[s].end_write(it)
            self.arrive(B["a_full"][warp])
            it += 4

    def mma_thread(self):
        NB, NST, B = self.NB, self.NST, self.bars

        def second_half(j):
            bj, sj, qj = j % NB, j % NST, j & 1
            yield ("wait", B["e1_done"][qj], (j >> 1) & 1)
            jj = j - NB
            if jj >= 0:
                yield ("wait", B["d_empty"][jj & 1], (jj >> 1) & 1)
            # G3: reads sDh[bj], writes accD[bj]; commit d_full
            self.mma(self.dur(0.05, 0.3), [self.sH[bj]], [self.accD[bj]], {self.sH[bj]: j, self.accD[bj]: j}, [B["d_full"][qj]])
            # G4 / G5: read sH/sDh[bj] and sA/sD[sj]; commit a_empty and h_free
            self.mma(self.dur(0.05, 0.4), [self.sH[bj], self.sAD[sj]], [], {self.sH[bj]: j, self.sAD[sj]: j},
                     [B["a_empty"][j & 3], B["h_free"][qj]])

        it = 0
        while it < self.ntiles:
            s, b = it % NST, it % NB
            yield ("wait", B["a_full"][it & 3], (it >> 2) & 1)
            if NB == 1 and it >= 1:
                yield from second_half(it - 1)
            # G1 / G2: read sA/sD[s], write acc1/accG[b]; commit hp_full
            self.mma(self.dur(0.05, 0.3), [self.sAD[s]], [self.acc[b]], {self.sAD[s]: it, self.acc[b]: it}, [B["hp_full"][it & 1]])
            if NB == 2 and it >= 1:
                yield from second_half(it - 1)
            yield ("sleep", self.dur(0.0, 0.05))
            it += 1
        if it >= 1:
            yield from second_half(it - 1)

    def epilogue(self, eg: int):
        NB, NST, B = self.NB, self.NST, self.bars
        it = eg
        while it < self.ntiles:
            b, s = it % NB, it % NST
            par = (it >> 1) & 1
            yield ("wait", B["a_full"][it & 3], (it >> 2) & 1)
            self.sAD[s].begin_read(it, f"E{eg} db3")           # conv3 bias gradient reads sD[s]
            yield ("sleep", self.dur(0.05, 0.3))
            self.sAD[s].end_read(it)
            yield ("wait", B["hp_full"][eg], par)
            pj = it - NB
            if pj >= 0:
                yield ("wait", B["h_free"][pj & 1], (pj >> 1) & 1)
            # E1: read acc1/accG[b], write sH/sDh[b]
            self.acc[b].begin_read(it, f"E{eg} E1")
            self.sH[b].begin_write(it)
            yield ("sleep", self.dur(0.5, 2.5))
            self.acc[b].end_read(it)
            self.sH[b].end_write(it)
            self.arrive(B["e1_done"][eg])
            yield ("wait", B["d_full"][eg], par)
            # E2: read accD[b]
            self.accD[b].begin_read(it, f"E{eg} E2")
            yield ("sleep", self.dur(0.3, 1.5))
            self.accD[b].end_read(it)
            self.done_tiles_e2 += 1
            self.arrive(B["d_empty"][eg])
            it += 2

    def run(self):
        roles = [self.loader(w) for w in range(4)] + [self.mma_thread(), self.epilogue(0), self.epilogue(1)]
        self.roles_alive = len(roles)
        for g in roles:
            self.at(0.0, lambda g=g: self.run_role(g))
        steps = 0
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            fn()
            steps += 1
            if steps > 200000 + 400 * self.ntiles:
                raise Hazard("runaway simulation")
        if self.roles_alive or self.waiting:
            stuck = [(b.name, p, b.phase) for _, b, p in self.waiting]
            raise Hazard(f"deadlock: {self.roles_alive} role(s) alive, waiting on {stuck}")
        if self.done_tiles_e2 != self.ntiles:
            raise Hazard(f"{self.done_tiles_e2} of {self.ntiles} tiles finished")


class SimL0(Sim):
    """`mlp_bwd_ws_kernel` as of round 2: THREE MMA-issuing threads (one per epilogue group for G1/G2/G3, one for the weight
    gradients G4/G5 of every tile in CTA order), accD double-buffered per group, and the epilogue order
    E1(k) -> e1_done -> E2(k-1).  Same index / parity formulas as the CUDA code; resources per group g: acc[g], sH[g],
    accD[g][b]."""

    def __init__(self, ntiles: int, seed: int):
        super().__init__(ntiles, 2, 4, seed)
        self.bars["d_full"] = [Barrier(f"d_full[{i}]", 1) for i in range(4)]
        self.bars["d_empty"] = [Barrier(f"d_empty[{i}]", 1) for i in range(4)]
        self.accD = [Resource(f"accD[{i}]") for i in range(4)]

    def issuer(self, g: int):
        B = self.bars
        k = 0
        while 2 * k + g < self.ntiles:
            it = 2 * k + g
            s, b = it & 3, k & 1
            yield ("wait", B["a_full"][s], (it >> 2) & 1)
            self.mma(self.dur(0.05, 0.3), [self.sAD[s]], [self.acc[g]], {self.sAD[s]: it, self.acc[g]: it}, [B["hp_full"][g]])
            yield ("wait", B["e1_done"][g], k & 1)
            if k >= 2:
                yield ("wait", B["d_empty"][g * 2 + b], ((k >> 1) - 1) & 1)
            self.mma(self.dur(0.05, 0.3), [self.sH[g]], [self.accD[g * 2 + b]], {self.sH[g]: it, self.accD[g * 2 + b]: it},
                     [B["d_full"][g * 2 + b]])
            yield ("sleep", self.dur(0.0, 0.05))
            k += 1

    def wgrad_issuer(self):
        B = self.bars
        for it in range(self.ntiles):
            g, s, k = it & 1, it & 3, it >> 1
            yield ("wait", B["a_full"][s], (it >> 2) & 1)
            yield ("wait", B["e1_done"][g], k & 1)
            self.mma(self.dur(0.05, 0.4), [self.sH[g], self.sAD[s]], [], {self.sH[g]: it, self.sAD[s]: it}, [B["a_empty"][s], B["h_free"][g]])
            yield ("sleep", self.dur(0.0, 0.05))

    def loader(self, warp: int):
        B = self.bars
        it, uses = warp, 0
        while it < self.ntiles:
            if uses >= 1:
                yield ("wait", B["a_empty"][warp], (uses - 1) & 1)
            self.sAD[warp].begin_write(it)
            yield ("sleep", self.dur(0.5, 3.0))
            self.sAD[warp].end_write(it)
            self.arrive(B["a_full"][warp])
            it += 4
            uses += 1

    def epilogue(self, eg: int):
        B = self.bars
        k = 0
        pending = None                                         # (tile, buffer) whose E2 is deferred

        def e2(tile, b):
            yield ("wait", B["d_full"][eg * 2 + b], ((tile >> 1) >> 1) & 1)
            self.accD[eg * 2 + b].begin_read(tile, f"E{eg} E2")
            yield ("sleep", self.dur(0.3, 1.5))
            self.accD[eg * 2 + b].end_read(tile)
            self.done_tiles_e2 += 1

        while 2 * k + eg < self.ntiles:
            it = 2 * k + eg
            s = it & 3
            yield ("wait", B["a_full"][s], (it >> 2) & 1)
            self.sAD[s].begin_read(it, f"E{eg} db3")
            yield ("sleep", self.dur(0.05, 0.3))
            self.sAD[s].end_read(it)
            yield ("wait", B["hp_full"][eg], k & 1)
            if k >= 1:
                yield ("wait", B["h_free"][eg], (k - 1) & 1)
                yield ("wait", B["d_full"][eg * 2 + ((k - 1) & 1)], ((k - 1) >> 1) & 1)
            self.acc[eg].begin_read(it, f"E{eg} E1")
            self.sH[eg].begin_write(it)
            yield ("sleep", self.dur(0.5, 2.5))
            self.acc[eg].end_read(it)
            self.sH[eg].end_write(it)
            self.arrive(B["e1_done"][eg])
            if pending is not None:
                yield from e2(*pending)
                self.arrive(B["d_empty"][eg * 2 + pending[1]])
            pending = (it, k & 1)
            k += 1
        if pending is not None:
            yield from e2(*pending)

    def run(self):
        roles = [self.loader(w) for w in range(4)] + [self.issuer(0), self.issuer(1), self.wgrad_issuer(), self.epilogue(0),
                                                        self.epilogue(1)]
        self.roles_alive = len(roles)
        for g in roles:
            self.at(0.0, lambda g=g: self.run_role(g))
        steps = 0
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            fn()
            steps += 1
            if steps > 200000 + 400 * self.ntiles:
                raise Hazard("runaway simulation")
        if self.roles_alive or self.waiting:
            stuck = [(b.name, p, b.phase) for _, b, p in self.waiting]
            raise Hazard(f"deadlock: {self.roles_alive} role(s) alive, waiting on {stuck}")
        if self.done_tiles_e2 != self.ntiles:
            raise Hazard(f"{self.done_tiles_e2} of {self.ntiles} tiles finished")


def check_l0(ntiles: int, seeds: int) -> None:
    for seed in range(seeds):
        SimL0(ntiles, seed).run()


class SimSplit(Sim):
    """`mlp_bwd_ws2_kernel<*, 1>` in column-split mode (round 2): one accumulator set, BOTH epilogue groups visit every tile and
    take half of the E1 / E2 columns each; e1_done / d_empty expect both groups, barriers are indexed by the tile parity."""

    def __init__(self, ntiles: int, seed: int):
        super().__init__(ntiles, 1, 2, seed)
        for name in ("e1_done", "d_empty"):
            self.bars[name] = [Barrier(f"{name}[{i}]", 2) for i in range(2)]
        self.sHh = [Resource("sH/sDh[lo]"), Resource("sH/sDh[hi]")]

    def mma_thread(self):
        B = self.bars

        def second_half(j):
            sj, qj = j % 2, j & 1
            yield ("wait", B["e1_done"][qj], (j >> 1) & 1)
            jj = j - 1
            if jj >= 0:
                yield ("wait", B["d_empty"][jj & 1], (jj >> 1) & 1)
            rd = {self.sHh[0]: j, self.sHh[1]: j, self.accD[0]: j}
            self.mma(self.dur(0.05, 0.3), [self.sHh[0], self.sHh[1]], [self.accD[0]], rd, [B["d_full"][qj]])
            rd2 = {self.sHh[0]: j, self.sHh[1]: j, self.sAD[sj]: j}
            self.mma(self.dur(0.05, 0.4), [self.sHh[0], self.sHh[1], self.sAD[sj]], [], rd2, [B["a_empty"][j & 3], B["h_free"][qj]])

        it = 0
        while it < self.ntiles:
            s = it % 2
            yield ("wait", B["a_full"][it & 3], (it >> 2) & 1)
            if it >= 1:
                yield from second_half(it - 1)
            self.mma(self.dur(0.05, 0.3), [self.sAD[s]], [self.acc[0]], {self.sAD[s]: it, self.acc[0]: it}, [B["hp_full"][it & 1]])
            yield ("sleep", self.dur(0.0, 0.05))
            it += 1
        if it >= 1:
            yield from second_half(it - 1)

    def epilogue(self, eg: int):
        B = self.bars
        for it in range(self.ntiles):
            s, q, par = it % 2, it & 1, (it >> 1) & 1
            yield ("wait", B["a_full"][it & 3], (it >> 2) & 1)
            if eg == 0:
                self.sAD[s].begin_read(it, "E0 db3")
                yield ("sleep", self.dur(0.05, 0.3))
                self.sAD[s].end_read(it)
            yield ("wait", B["hp_full"][q], par)
            pj = it - 1
            if pj >= 0:
                yield ("wait", B["h_free"][pj & 1], (pj >> 1) & 1)
            self.acc[0].begin_read(it, f"E{eg} E1")
            self.sHh[eg].begin_write(it)
            yield ("sleep", self.dur(0.3, 1.5))
            self.acc[0].end_read(it)
            self.sHh[eg].end_write(it)
            self.arrive(B["e1_done"][q])
            yield ("wait", B["d_full"][q], par)
            self.accD[0].begin_read(it, f"E{eg} E2")
            yield ("sleep", self.dur(0.2, 1.0))
            self.accD[0].end_read(it)
            if eg == 0:
                self.done_tiles_e2 += 1
            self.arrive(B["d_empty"][q])


def check_split(ntiles: int, seeds: int) -> None:
    for seed in range(seeds):
        SimSplit(ntiles, seed).run()


class SimFwd(Sim):
    """`mlp_fused_kernel` (csrc/mednext_fwd.cu) as of round 2: four A stages (loader warp w <-> stage w, tiles it == w mod 4), one
    MMA-issuing thread per epilogue group, acc2 with NB2 buffers per group; with NB2 == 2 a group runs epilogue 2 of tile k-1
    after epilogue 1 of tile k.  Resources per group g: acc1[g], sH[g], acc2[g][b]."""

    def __init__(self, ntiles: int, nb2: int, seed: int, has_rc: bool = False):
        super().__init__(ntiles, 2, 4, seed)
        self.nb2, self.has_rc = nb2, has_rc
        self.bars = {
            "a_full": [Barrier(f"a_full[{i}]", 1) for i in range(4)],
            "a_empty": [Barrier(f"a_empty[{i}]", 1) for i in range(4)],
            "acc1_full": [Barrier(f"acc1_full[{i}]", 1) for i in range(2)],
            "h_full": [Barrier(f"h_full[{i}]", 1) for i in range(2)],
            "acc2_full": [Barrier(f"acc2_full[{i}]", 1) for i in range(4)],
            "acc2_empty": [Barrier(f"acc2_empty[{i}]", 1) for i in range(4)],
        }
        self.sA = [Resource(f"sA[{i}]") for i in range(4)]
        self.acc1 = [Resource(f"acc1[{i}]") for i in range(2)]
        self.sHf = [Resource(f"sH[{i}]") for i in range(2)]
        self.acc2 = [Resource(f"acc2[{i}]") for i in range(4)]

    def loader(self, warp: int):
        B = self.bars
        it, uses = warp, 0
        while it < self.ntiles:
            if uses >= 1:
                yield ("wait", B["a_empty"][warp], (uses - 1) & 1)
            self.sA[warp].begin_write(it)
            yield ("sleep", self.dur(0.3, 2.0))
            self.sA[warp].end_write(it)
            self.arrive(B["a_full"][warp])
            it += 4
            uses += 1

    def issuer(self, g: int):
        B, NB2 = self.bars, self.nb2
        k = 0
        while 2 * k + g < self.ntiles:
            it = 2 * k + g
            s = it % 4
            yield ("wait", B["a_full"][s], (it // 4) & 1)
            commits = [B["acc1_full"][g]] + ([] if self.has_rc else [B["a_empty"][s]])
            self.mma(self.dur(0.05, 0.3), [self.sA[s]], [self.acc1[g]], {self.sA[s]: it, self.acc1[g]: it}, commits)
            yield ("wait", B["h_full"][g], k & 1)
            b = (k & 1) if NB2 == 2 else 0
            prev = k - NB2
            if prev >= 0:
                yield ("wait", B["acc2_empty"][g * 2 + b], (prev // NB2) & 1)
            reads = [self.sHf[g]] + ([self.sA[s]] if self.has_rc else [])
            tiles = {self.sHf[g]: it, self.acc2[g * 2 + b]: it}
            if self.has_rc:
                tiles[self.sA[s]] = it
            commits = [B["acc2_full"][g * 2 + b]] + ([B["a_empty"][s]] if self.has_rc else [])
            self.mma(self.dur(0.05, 0.3), reads, [self.acc2[g * 2 + b]], tiles, commits)
            yield ("sleep", self.dur(0.0, 0.05))
            k += 1

    def epilogue(self, eg: int):
        B, NB2 = self.bars, self.nb2
        pipe = NB2 == 2
        k, pending = 0, None

        def e2(tile, b):
            self.acc2[eg * 2 + b].begin_read(tile, f"E{eg} epi2")
            yield ("sleep", self.dur(0.2, 1.0))
            self.acc2[eg * 2 + b].end_read(tile)
            self.done_tiles_e2 += 1
            self.arrive(B["acc2_empty"][eg * 2 + b])

        while 2 * k + eg < self.ntiles:
            it = 2 * k + eg
            yield ("wait", B["acc1_full"][eg], k & 1)
            if pipe and k >= 1:
                yield ("wait", B["acc2_full"][eg * 2 + ((k - 1) & 1)], ((k - 1) >> 1) & 1)
            self.acc1[eg].begin_read(it, f"E{eg} epi1")
            self.sHf[eg].begin_write(it)
            yield ("sleep", self.dur(0.4, 2.0))
            self.acc1[eg].end_read(it)
            self.sHf[eg].end_write(it)
            self.arrive(B["h_full"][eg])
            if pipe:
                if pending is not None:
                    yield from e2(*pending)
                pending = (it, k & 1)
            else:
                yield ("wait", B["acc2_full"][eg * 2], k & 1)
                yield from e2(it, 0)
            k += 1
        if pipe and pending is not None:
            yield ("wait", B["acc2_full"][eg * 2 + pending[1]], ((pending[0] >> 1) >> 1) & 1)
            yield from e2(*pending)

    def run(self):
        roles = [self.loader(w) for w in range(4)] + [self.issuer(0), self.issuer(1), self.epilogue(0), self.epilogue(1)]
        self.roles_alive = len(roles)
        for g in roles:
            self.at(0.0, lambda g=g: self.run_role(g))
        steps = 0
        while self.events:
            t, _, fn = heapq.heappop(self.events)
            self.now = t
            fn()
            steps += 1
            if steps > 200000 + 400 * self.ntiles:
                raise Hazard("runaway simulation")
        if self.roles_alive or self.waiting:
            stuck = [(b.name, p, b.phase) for _, b, p in self.waiting]
            raise Hazard(f"deadlock: {self.roles_alive} role(s) alive, waiting on {stuck}")
        if self.done_tiles_e2 != self.ntiles:
            raise Hazard(f"{self.done_tiles_e2} of {self.ntiles} tiles finished")


def check_fwd(ntiles: int, nb2: int, seeds: int, has_rc: bool = False) -> None:
    for seed in range(seeds):
        SimFwd(ntiles, nb2, seed, has_rc).run()


def check(ntiles: int, NB: int, NST: int, seeds: int) -> None:
    for seed in range(seeds):
        Sim(ntiles, NB, NST, seed).run()


CONFIGS = [(2, 4), (1, 2), (1, 3), (1, 4)]      # (NB, NST): double-buffered / single-buffered accumulators with 2..4 operand stages


def main(seeds: int = 200) -> int:
    for NB, NST in CONFIGS:
        for ntiles in (1, 2, 3, 4, 5, 7, 8, 9, 16, 33):
            check(ntiles, NB, NST, seeds)
        print(f"NB={NB} NST={NST}: ok ({seeds} schedules x 10 tile counts)")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 200))

"""Barrier protocol of the warp-specialised backward kernels (csrc/mednext_bwd.cu: mlp_bwd_ws_kernel, mlp_bwd_ws2_kernel),
checked on CPU with the discrete-event model in tools/ws_protocol_model.py: no deadlock, no parity aliasing, no buffer
overwritten while it is still read — for both accumulator configurations, many tile counts and random timings."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import ws_protocol_model as M  # noqa: E402


@pytest.mark.parametrize("NB,NST", M.CONFIGS)
@pytest.mark.parametrize("ntiles", [1, 2, 3, 4, 5, 8, 9, 21])
def test_protocol_has_no_hazard(NB, NST, ntiles):
    M.check(ntiles, NB, NST, seeds=40)


@pytest.mark.parametrize("ntiles", [1, 2, 3, 4, 5, 6, 8, 9, 13, 21])
def test_level0_three_issuer_protocol_has_no_hazard(ntiles):
    """mlp_bwd_ws_kernel (round 2): per-group issuers + weight-gradient issuer, deferred E2, accD double-buffered per group."""
    M.check_l0(ntiles, seeds=40)


@pytest.mark.parametrize("ntiles", [1, 2, 3, 4, 5, 8, 9, 21])
def test_column_split_protocol_has_no_hazard(ntiles):
    """mlp_bwd_ws2_kernel<*, 1> with both epilogue groups on every tile (e1_done / d_empty count both groups)."""
    M.check_split(ntiles, seeds=40)


@pytest.mark.parametrize("nb2", [1, 2])
@pytest.mark.parametrize("has_rc", [False, True])
@pytest.mark.parametrize("ntiles", [1, 2, 3, 5, 8, 9, 21])
def test_forward_two_issuer_protocol_has_no_hazard(ntiles, nb2, has_rc):
    """mlp_fused_kernel (round 2): one MMA issuer per epilogue group, acc2 double-buffered + deferred epilogue 2 (nb2 = 2) or the
    round-1 order (nb2 = 1), with and without the res-conv GEMM holding the A stage until GEMM2 retires."""
    M.check_fwd(ntiles, nb2, seeds=30, has_rc=has_rc)


def test_level0_model_flags_a_missing_g3_wait():
    """E1(k) must see G3(k-1) retired (another issuer than the one that commits h_free): without that wait the model reports
    sDh overwritten while G3 still reads it (or a wrong-tile read)."""

    class _NoG3Wait(M.SimL0):
        def epilogue(self, eg):
            for op in super().epilogue(eg):
                if op[0] == "wait" and op[1].name.startswith("d_full") and getattr(self, "_skip", True):
                    continue
                yield op

    bad = 0
    for seed in range(60):
        try:
            _NoG3Wait(9, seed).run()
        except M.Hazard:
            bad += 1
    assert bad > 0


class _BufferIndexed(M.Sim):
    """The first draft of the generalised kernel indexed the accumulator barriers by BUFFER (it % NB): with one buffer the
    two epilogue groups then share a barrier and each skips every other completion.  The model must flag it."""

    def mma_thread(self):
        NB, NST, B = self.NB, self.NST, self.bars

        def second_half(j):
            bj, sj, u = j % NB, j % NST, j // NB
            yield ("wait", B["e1_done"][bj], u & 1)
            if j >= NB:
                yield ("wait", B["d_empty"][bj], (u - 1) & 1)
            self.mma(0.1, [self.sH[bj]], [self.accD[bj]], {self.sH[bj]: j, self.accD[bj]: j}, [B["d_full"][bj]])
            self.mma(0.1, [self.sH[bj], self.sAD[sj]], [], {self.sH[bj]: j, self.sAD[sj]: j}, [B["a_empty"][j & 3], B["h_free"][bj]])

        it = 0
        while it < self.ntiles:
            s, b = it % NST, it % NB
            yield ("wait", B["a_full"][it & 3], (it >> 2) & 1)
            if NB == 1 and it >= 1:
                yield from second_half(it - 1)
            self.mma(0.1, [self.sAD[s]], [self.acc[b]], {self.sAD[s]: it, self.acc[b]: it}, [B["hp_full"][b]])
            if NB == 2 and it >= 1:
                yield from second_half(it - 1)
            yield ("sleep", 0.01)
            it += 1
        if it >= 1:
            yield from second_half(it - 1)

    def epilogue(self, eg):
        NB, NST, B = self.NB, self.NST, self.bars
        it = eg
        while it < self.ntiles:
            b, s, u = it % NB, it % NST, it // NB
            yield ("wait", B["a_full"][it & 3], (it >> 2) & 1)
            yield ("wait", B["hp_full"][b], u & 1)
            if u >= 1:
                yield ("wait", B["h_free"][b], (u - 1) & 1)
            self.acc[b].begin_read(it, "E1")
            self.sH[b].begin_write(it)
            yield ("sleep", self.dur(0.5, 2.5))
            self.acc[b].end_read(it)
            self.sH[b].end_write(it)
            self.arrive(B["e1_done"][b])
            yield ("wait", B["d_full"][b], u & 1)
            self.accD[b].begin_read(it, "E2")
            yield ("sleep", self.dur(0.3, 1.5))
            self.accD[b].end_read(it)
            self.done_tiles_e2 += 1
            self.arrive(B["d_empty"][b])
            it += 2


def test_model_flags_the_buffer_indexed_draft():
    flagged = 0
    for seed in range(40):
        try:
            _BufferIndexed(16, 1, 2, seed).run()
        except M.Hazard:
            flagged += 1
    assert flagged > 0                       # parity aliasing shows up as a wrong-tile read or a deadlock
    for seed in range(10):                   # with two buffers tile-indexed and buffer-indexed barriers coincide
        _BufferIndexed(16, 2, 4, seed).run()

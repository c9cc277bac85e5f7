"""``pcb_comm`` — the library's own NCCL communicator (C ABI: ``pcb_comm_init`` / ``pcb_grad_allreduce`` /
``pcb_sw_exchange_overlap``, ``include/pcb200.h``), bound from Python.

The Python package's default exchange is ``torch.distributed`` (``training/ddp.py``, ``inference/sharded.py``); this handle is
the SAME two exchange steps through the C ABI, i.e. what a host without ``torch.distributed`` binds, and what
``FlatGradArena.allreduce_sum(comm=...)`` / ``exchange_overlaps(..., comm=...)`` use when a :class:`NativeComm` is passed.
There is no host fallback: without a CUDA device or an NCCL in the process the constructor raises.
"""

from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L

ID_BYTES = 128


def unique_id() -> bytes:
    """``ncclGetUniqueId`` through the library (rank 0 calls it and hands the 128 bytes to every rank)."""
    buf = ctypes.create_string_buffer(ID_BYTES)
    L.check(L.lib().pcb_comm_unique_id(buf), "pcb_comm_unique_id")
    return buf.raw


class NativeComm:
    """One communicator per rank process, tied to the CUDA device that is current at construction.

    ``NativeComm(uid, rank, world)`` is collective over all ranks.  :meth:`from_process_group` bootstraps the id over an
    existing ``torch.distributed`` group (any backend — only the 128 id bytes travel through it)."""

    def __init__(self, uid: bytes, rank: int, world: int, device: Optional[torch.device] = None) -> None:
        if len(uid) != ID_BYTES:
            raise ValueError(f"NativeComm: the unique id must be {ID_BYTES} bytes, got {len(uid)}")
        if not torch.cuda.is_available():
            raise RuntimeError("pcb200: NativeComm needs a CUDA device (there is no host path for the exchange)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().pcb_comm_init(ctypes.c_char_p(uid), int(rank), int(world), ctypes.byref(self._h)), "pcb_comm_init")
        self.rank, self.world = int(rank), int(world)

    @classmethod
    def from_process_group(cls, group=None, device: Optional[torch.device] = None) -> "NativeComm":
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return cls(unique_id(), 0, 1, device)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(box[0], rank, world, device)

    def _check(self, t: torch.Tensor, what: str) -> None:
        if self._h is None or not self._h.value:
            raise RuntimeError("pcb200: NativeComm is closed")
        L.require_device(t, what)
        if t.device != self.device:
            raise ValueError(f"{what}: tensor on {t.device}, communicator on {self.device}")
        if not t.is_contiguous():
            raise ValueError(f"{what}: tensors must be contiguous")

    def allreduce_(self, t: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        """in place: ``t = scale * sum over ranks``, enqueued on the current stream"""
        self._check(t, "NativeComm.allreduce_")
        L.check(L.lib().pcb_grad_allreduce(self._h, L.ptr(t), ctypes.c_int64(t.numel()), L.dtype_code(t.dtype),
                                           ctypes.c_float(float(scale)), L.stream_ptr(t.device)), "pcb_grad_allreduce")
        return t

    def exchange(self, sends: Sequence[Tuple[torch.Tensor, int]], recvs: Sequence[Tuple[torch.Tensor, int]]) -> None:
        """one grouped send/recv: ``sends`` = [(tensor, peer)], ``recvs`` = [(buffer, peer)] (same dtype everywhere)"""
        if not sends and not recvs:
            return
        every = [t for t, _ in list(sends) + list(recvs)]
        for t in every:
            self._check(t, "NativeComm.exchange")
        if any(t.dtype != every[0].dtype for t in every):
            raise ValueError("NativeComm.exchange: all messages must share one dtype")

        def pack(items):
            n = len(items)
            return ((ctypes.c_void_p * n)(*[t.data_ptr() for t, _ in items]),
                    (ctypes.c_int64 * n)(*[t.numel() for t, _ in items]), (ctypes.c_int * n)(*[int(p) for _, p in items]), n)

        sb, sn, sp, ns = pack(list(sends))
        rb, rn, rp, nr = pack(list(recvs))
        L.check(L.lib().pcb_sw_exchange_overlap(self._h, ns, sb, sn, sp, nr, rb, rn, rp, L.dtype_code(every[0].dtype),
                                                L.stream_ptr(self.device)), "pcb_sw_exchange_overlap")

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            L.lib().pcb_comm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):          # best effort; explicit close() is the documented way
        try:
            self.close()
        except Exception:
            pass


def nccl_version() -> int:
    """NCCL_VERSION_CODE of the library NCCL calls are bound to (-1: none in the process / on the loader path)"""
    return int(L.lib().pcb_comm_nccl_version())

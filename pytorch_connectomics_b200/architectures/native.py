"""``pcb_net`` binding: a MedNeXt's forward plan handed to the library (``include/pcb200.h``, csrc/net_runtime.cu).

The ``nn.Module`` tree stays the parameter container (and the training path); for inference its blocks are flattened
into ``pcb_block_desc`` records carrying device pointers to the kernel-layout weights, after which one
``pcb_net_forward`` enqueues the whole network and ``pcb_sw_run`` the whole sliding-window tile loop
(``connectomics/inference/window.py:563-683``) without returning to Python between kernels.
"""

from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from .. import _lib as L
from . import _mednext_ops as ops


class BlockDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("C", ctypes.c_int32), ("H", ctypes.c_int32), ("Co", ctypes.c_int32),
                ("k", ctypes.c_int32), ("do_res", ctypes.c_int32), ("norm", ctypes.c_int32), ("skip_from", ctypes.c_int32),
                ("w1", ctypes.c_void_p), ("b1", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
                ("w2", ctypes.c_void_p), ("b2", ctypes.c_void_p), ("w3", ctypes.c_void_p), ("b3", ctypes.c_void_p),
                ("wr", ctypes.c_void_p), ("br", ctypes.c_void_p)]


class HeadDesc(ctypes.Structure):
    _fields_ = [("from_block", ctypes.c_int32), ("ncls", ctypes.c_int32), ("w", ctypes.c_void_p), ("b", ctypes.c_void_p)]


def native_eligible(model) -> Optional[str]:
    """``None`` when the trunk can run on the native plan, else the reason it cannot."""
    from .mednext import MedNeXt, MedNeXtBlock
    if not isinstance(model, MedNeXt):
        return "not a pcb200 MedNeXt"
    for m in model.modules():
        if isinstance(m, MedNeXtBlock) and (m.norm_type != "group" or m.grn or m.dim != "3d"):
            return "blocks with LayerNorm / GRN / 2-D run through the module path"
    stages = [model.enc_block_0, model.enc_block_1, model.enc_block_2, model.enc_block_3, model.bottleneck,
              model.dec_block_3, model.dec_block_2, model.dec_block_1, model.dec_block_0]
    if any(len(s) == 0 for s in stages):
        return "a stage without blocks"
    return None


class NativeMedNeXt:
    """Owns a ``pcb_net`` built from ``model`` (a ``architectures.mednext.MedNeXt``) on ``model``'s device."""

    def __init__(self, model) -> None:
        why = native_eligible(model)
        if why is not None:
            raise ValueError(f"pcb200 native plan: {why}")
        self.model = model
        self.device = model.stem.weight.device
        L.require_device(model.stem.weight, "pcb_net")
        self._keep: List[torch.Tensor] = []        # kernel-layout weights referenced by the plan
        self._sig = self._signature()
        blocks: List[BlockDesc] = []
        ends = {}

        def pk(p, kind):
            t = ops.packed(p, kind)
            self._keep.append(t)
            return t.data_ptr()

        def add(blk, skip_from=-1):
            rc = getattr(blk, "res_conv", None)
            mode = blk._dw_mode
            d = BlockDesc(kind=mode, C=blk.conv1.in_channels, H=blk.conv2.out_channels, Co=blk.conv3.out_channels,
                          k=blk.conv1.kernel_size[0], do_res=1 if (mode == L.DW_SAME and blk.do_res) else 0, norm=0,
                          skip_from=skip_from, w1=pk(blk.conv1.weight, "dw"), b1=pk(blk.conv1.bias, "f32"),
                          gamma=pk(blk.norm.weight, "f32"), beta=pk(blk.norm.bias, "f32"), w2=pk(blk.conv2.weight, "pw"),
                          b2=pk(blk.conv2.bias, "f32"), w3=pk(blk.conv3.weight, "pw"), b3=pk(blk.conv3.bias, "f32"),
                          wr=pk(rc.weight, "pw_t" if mode == L.DW_UP else "pw") if rc is not None else None,
                          br=pk(rc.bias, "f32") if rc is not None else None)
            blocks.append(d)
            return len(blocks) - 1

        def stage(seq, name):
            for b in seq:
                ends[name] = add(b)

        m = model
        stage(m.enc_block_0, "r0"); add(m.down_0)
        stage(m.enc_block_1, "r1"); add(m.down_1)
        stage(m.enc_block_2, "r2"); add(m.down_2)
        stage(m.enc_block_3, "r3"); add(m.down_3)
        stage(m.bottleneck, "b")
        add(m.up_3, ends["r3"]); stage(m.dec_block_3, "d3")
        add(m.up_2, ends["r2"]); stage(m.dec_block_2, "d2")
        add(m.up_1, ends["r1"]); stage(m.dec_block_1, "d1")
        add(m.up_0, ends["r0"]); stage(m.dec_block_0, "f0")
        heads = [(m.out_0, ends["f0"])]
        if m.do_ds:
            heads += [(m.out_1, ends["d1"]), (m.out_2, ends["d2"]), (m.out_3, ends["d3"]), (m.out_4, ends["b"])]
        hd = (HeadDesc * len(heads))(*[HeadDesc(from_block=fb, ncls=ob.conv_out.out_channels,
                                                w=pk(ob.conv_out.weight, "head"), b=pk(ob.conv_out.bias, "f32"))
                                       for ob, fb in heads])
        self.head_channels = [int(ob.conv_out.out_channels) for ob, _ in heads]
        self.head_levels = [0, 1, 2, 3, 4][:len(heads)]
        bd = (BlockDesc * len(blocks))(*blocks)
        handle = ctypes.c_void_p()
        L.check(L.lib().pcb_net_create(ctypes.c_int32(m.stem.in_channels), ctypes.c_int32(m.stem.out_channels),
                                       ctypes.c_void_p(pk(m.stem.weight, "f32")), ctypes.c_void_p(pk(m.stem.bias, "f32")),
                                       bd, ctypes.c_int32(len(blocks)), hd, ctypes.c_int32(len(heads)), ctypes.byref(handle)),
                "pcb_net_create")
        self.handle = handle
        self.in_channels = int(m.stem.in_channels)
        self._ws: Optional[torch.Tensor] = None

    def _signature(self):
        return tuple((id(p), p._version, p.data_ptr()) for p in self.model.parameters()) + (L.PARAM_EPOCH[0],)

    def stale(self) -> bool:
        """weights changed (optimizer step, load_state_dict, .to()) since the plan was built"""
        return self._sig != self._signature()

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                L.lib().pcb_net_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    def workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), device=self.device, dtype=torch.uint8)
        return self._ws

    def forward(self, x: torch.Tensor, out_dtype: Optional[torch.dtype] = None, heads: Optional[Sequence[int]] = None):
        """``x`` NCDHW on the plan's device -> list of head outputs (``None`` for heads not asked for)."""
        L.require_device(x, "pcb_net_forward")
        if x.dim() != 5 or int(x.shape[1]) != self.in_channels:
            raise ValueError(f"MedNeXt expects (B, {self.in_channels}, D, H, W); got shape {tuple(x.shape)}")
        if any(int(s) % 16 for s in x.shape[2:]):
            raise ValueError(f"MedNeXt input spatial size must be divisible by 16, got {tuple(x.shape[2:])}")
        x = x.contiguous()
        n, size = int(x.shape[0]), [int(s) for s in x.shape[2:]]
        odt = out_dtype or (x.dtype if x.dtype in (torch.float16, torch.bfloat16, torch.float32) else torch.float32)
        want = list(range(len(self.head_channels))) if heads is None else list(heads)
        outs = [torch.empty((n, c, *[s >> lv for s in size]), device=x.device, dtype=odt) if h in want else None
                for h, (c, lv) in enumerate(zip(self.head_channels, self.head_levels))]
        lib = L.lib()
        nbytes = int(lib.pcb_net_workspace_bytes(self.handle, ctypes.c_int64(n), L.i64x(size)))
        if nbytes < 0:
            L.check(-1, "pcb_net_workspace_bytes")
        ws = self.workspace(nbytes)
        ptrs = (ctypes.c_void_p * len(outs))(*[None if o is None else o.data_ptr() for o in outs])
        with torch.cuda.device(x.device):
            L.check(lib.pcb_net_forward(self.handle, L.ptr(x), L.dtype_code(x.dtype), ctypes.c_int64(n), L.i64x(size), ptrs,
                                        L.dtype_code(odt), L.ptr(ws), ctypes.c_int64(ws.numel()), L.stream_ptr(x.device)),
                    "pcb_net_forward")
        return outs

    def sw_run(self, vol: torch.Tensor, roi, starts, wmap: torch.Tensor, value: torch.Tensor, weight: torch.Tensor, *,
               padding_mode: str, cval: float, sw_batch: int, use_graph: bool = True, head: int = 0) -> None:
        """``pcb_sw_run``: every window of ``starts`` (grid order) cropped from ``vol`` [1,C,D,H,W], predicted and blended
        into ``value`` / ``weight`` (not zeroed, not normalised)."""
        lib = L.lib()
        roi = [int(v) for v in roi]
        image = [int(v) for v in vol.shape[2:]]
        acc = [int(v) for v in value.shape[2:]]
        flat = [int(c) for s in starts for c in s]
        n = len(starts)
        bs = max(1, min(int(sw_batch), 16, 8))          # the fused MedNeXt kernels serve up to 8 samples per launch
        vdt, adt = L.dtype_code(vol.dtype), L.dtype_code(value.dtype)
        nbytes = int(lib.pcb_sw_run_workspace_bytes(self.handle, head, L.i64x(roi), bs, ctypes.c_int64(n), vdt, adt))
        if nbytes < 0:
            L.check(-1, "pcb_sw_run_workspace_bytes")
        ws = self.workspace(nbytes)
        with torch.cuda.device(vol.device):
            L.check(lib.pcb_sw_run(self.handle, head, L.ptr(vol), vdt, L.i64x(image), L.i64x(roi), L.PAD[padding_mode],
                                   ctypes.c_double(cval), bs, L.i64x(flat), ctypes.c_int64(n), L.ptr(wmap), L.ptr(value),
                                   L.ptr(weight), adt, L.i64x(acc), L.ptr(ws), ctypes.c_int64(ws.numel()),
                                   1 if use_graph else 0, L.stream_ptr(vol.device)), "pcb_sw_run")

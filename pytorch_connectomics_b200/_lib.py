"""ctypes binding of ``csrc/libpcb200.so`` (the C ABI declared in ``include/pcb200.h``).

The product path has NO CPU fallback: if the shared library is missing, or a CUDA op is asked to
run without a Blackwell device, a ``RuntimeError`` is raised.  Bad arguments surface as
``ValueError`` (status PCB_ERR_INVALID) with the library's message, mirroring the reference's
error behaviour for the same misuse.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libpcb200.so")
SOURCES = ["pcb_api.cu", "sw_kernels.cu", "mednext_fwd.cu", "mednext_bwd.cu", "dense_conv.cu", "deep_mlp.cu", "layernorm.cu", "tta_kernels.cu",
           "optim_kernels.cu", "net_runtime.cu", "comm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

PCB_F32, PCB_F16, PCB_BF16 = 0, 1, 2
BLEND = {"constant": 0, "bump": 1, "distance": 2}
PAD = {"constant": 0, "reflect": 1, "replicate": 2, "circular": 3}
GRID_EAGER, GRID_LAZY, GRID_LAZY_SNAP = 0, 1, 2
DW_SAME, DW_DOWN, DW_UP = 0, 1, 2

_DTYPES = {torch.float32: PCB_F32, torch.float16: PCB_F16, torch.bfloat16: PCB_BF16}

_lib: Optional[ctypes.CDLL] = None

# Bumped by anything that rewrites parameter storage behind autograd's back (the fused optimizer kernel updates the flat
# parameter arena through a raw pointer, so ``Parameter._version`` does not move): kernel-layout weight caches
# (``architectures/_mednext_ops.packed``) compare it and repack.
PARAM_EPOCH = [0]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources into ``csrc/libpcb200.so`` with nvcc for sm_100a (in-tree): one object per source,
    compiled in parallel and only when stale, then one link step."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, "pcb_common.cuh"), os.path.join(_HERE, "..", "include", "pcb200.h")]
    hdr_t = max(os.path.getmtime(h) for h in hdrs if os.path.exists(h))
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def obj_of(src):
        return os.path.join(objdir, os.path.basename(src)[:-3] + ".o")

    def stale(src):
        o = obj_of(src)
        return force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(src), hdr_t)

    todo = [s for s in srcs if stale(s)]

    def compile_one(src):
        cmd = [nvcc] + flags + ["-c", "-o", obj_of(src), src]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=CSRC)

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            list(ex.map(compile_one, todo))
    objs = [obj_of(s) for s in srcs]
    if todo or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs + ["-ldl"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"pcb200: CUDA extension {LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback for this path.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pcb_last_error.restype = ctypes.c_char_p
        _lib.pcb_version.restype = ctypes.c_int
        _lib.pcb_launch_count.restype = ctypes.c_int64
        _lib.pcb_tn_workspace_floats.restype = ctypes.c_int64
        _lib.pcb_mlp_bwd_fused_workspace_floats.restype = ctypes.c_int64
        _lib.pcb_mlp_fwd_deep_workspace.restype = ctypes.c_int64
        _lib.pcb_net_workspace_bytes.restype = ctypes.c_int64
        _lib.pcb_sw_run_workspace_bytes.restype = ctypes.c_int64
        _lib.pcb_net_destroy.restype = None
        _lib.pcb_net_destroy.argtypes = [ctypes.c_void_p]
        _lib.pcb_comm_destroy.restype = None
        _lib.pcb_comm_destroy.argtypes = [ctypes.c_void_p]
        _lib.pcb_comm_rank.argtypes = [ctypes.c_void_p]
        _lib.pcb_comm_world.argtypes = [ctypes.c_void_p]
    return _lib


def launch_count() -> int:
    return int(lib().pcb_launch_count())


# ---- optional per-op CUDA-event timing (bench.py roofline leg); None = disabled, zero overhead
_PROF = None


class prof:
    """``with prof("mlp_fwd:C32..."):`` records start/stop events on the current stream when enabled."""

    __slots__ = ("name", "e0")

    def __init__(self, name: str):
        self.name = name
        self.e0 = None

    def __enter__(self):
        if _PROF is not None and (not _PROF["want"] or self.name in _PROF["want"]):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _PROF["ev"].setdefault(self.name, []).append((self.e0, e1))
        return False


def prof_start(want=()):
    global _PROF
    _PROF = {"want": set(want), "ev": {}}


def prof_stop():
    """Returns {name: [ms, ...]} (synchronises)."""
    global _PROF
    p, _PROF = _PROF, None
    torch.cuda.synchronize()
    return {k: [a.elapsed_time(b) for a, b in v] for k, v in (p["ev"] if p else {}).items()}


def check(status: int, what: str = "") -> None:
    if status == 0:
        return
    msg = lib().pcb_last_error().decode("utf-8", "replace")
    if status == -1:
        raise ValueError(msg or f"pcb200: invalid argument in {what}")
    raise RuntimeError(f"pcb200 {what} failed (status {status}): {msg}")


def require_device(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"pcb200: {what} needs a CUDA tensor on a B200 (sm_100a); got device={t.device}. "
            "There is no CPU fallback for this path.")


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise ValueError(f"pcb200: unsupported dtype {dt}; expected float32/float16/bfloat16") from None


def i64x(vals: Sequence[int]):
    return (ctypes.c_int64 * len(vals))(*[int(v) for v in vals])


def f64x(vals: Sequence[float]):
    return (ctypes.c_double * len(vals))(*[float(v) for v in vals])


def ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def exported_symbols():
    """Names declared in include/pcb200.h (parsed) — used by the CPU-side load test."""
    import re
    hdr = os.path.join(_HERE, "..", "include", "pcb200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(pcb_[a-z0-9_]+)\s*\(", txt)))

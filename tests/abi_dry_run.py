"""TEST INFRASTRUCTURE — a dry-run double of ``libpcb200.so`` for the CPU suite.  ``include/pcb200.h`` is parsed into
prototypes; the double answers every ``pcb_*`` call with success WITHOUT computing anything, after checking what ctypes
would silently get wrong on a real call: the argument COUNT, pointers where the prototype has pointers (``c_void_p``, a
ctypes array, ``None``/``byref`` — never a bare Python number), ``c_float`` / ``c_double`` exactly where the prototype says
``float`` / ``double`` (ctypes cannot convert a Python float and passes a mismatched width as garbage), integers elsewhere.
With it the host code that sits between torch and the C ABI — autograd wrappers, shape bookkeeping, workspace sizing,
gradient reshapes — runs end to end on CPU tensors, so code written without GPU access meets its first B200 run free of
Python-level and call-signature errors.  Values are meaningless (outputs are uninitialised); only shapes, dtypes and the
calls are checked.  Never reachable from the product package."""

from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

from conftest import ROOT

_INT_TYPES = (ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_long, ctypes.c_longlong, ctypes.c_uint, ctypes.c_uint32,
              ctypes.c_uint64, ctypes.c_ulong, ctypes.c_size_t, ctypes.c_bool)


def parse_header() -> Dict[str, Tuple[str, List[str]]]:
    """{name: (return type, [parameter kind])} with kinds 'ptr' | 'float' | 'double' | 'int'"""
    txt = open(os.path.join(ROOT, "include", "pcb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", " ", txt)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(pcb_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", txt, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        kinds = []
        if params not in ("", "void"):
            for prm in params.split(","):
                prm = prm.strip()
                if "*" in prm or "[" in prm:
                    kinds.append("ptr")
                elif re.search(r"\bfloat\b", prm):
                    kinds.append("float")
                elif re.search(r"\bdouble\b", prm):
                    kinds.append("double")
                else:
                    kinds.append("int")
        protos[name] = (ret, kinds)
    return protos


class _Fn:
    def __init__(self, lib, name, ret, kinds):
        self.lib, self.name, self.ret, self.kinds = lib, name, ret, kinds
        self.restype = ctypes.c_int
        self.argtypes = None

    def __call__(self, *args):
        lib, name = self.lib, self.name
        lib.calls.append(name)
        assert len(args) == len(self.kinds), f"{name}: {len(args)} arguments for a prototype with {len(self.kinds)}"
        for i, (a, kind) in enumerate(zip(args, self.kinds)):
            where = f"{name} argument {i} ({kind})"
            if kind == "ptr":
                ok = a is None or isinstance(a, (ctypes.c_void_p, ctypes.c_char_p, ctypes.Array, bytes, ctypes._Pointer)) \
                    or type(a).__name__ == "CArgObject" or (self.argtypes is not None and isinstance(a, int))
                assert ok, f"{where}: got {type(a).__name__} {a!r}"
            elif kind == "float":
                assert isinstance(a, ctypes.c_float), f"{where}: got {type(a).__name__}"
            elif kind == "double":
                assert isinstance(a, ctypes.c_double), f"{where}: got {type(a).__name__}"
            else:
                assert (isinstance(a, int) and not isinstance(a, bool)) or isinstance(a, (bool,) + _INT_TYPES), \
                    f"{where}: got {type(a).__name__} {a!r}"
        if name in lib.effects:
            lib.effects[name](*args)
        if name in lib.returns:
            return lib.returns[name]
        if name == "pcb_last_error":
            return b""
        if "*" in self.ret or self.ret.startswith("void"):
            return None
        if re.search(r"workspace|_floats$|_bytes$", name):
            return 64
        return 0


class DryRunLib:
    """``returns``: per-function overrides of the answer (e.g. ``{"pcb_mlp_bwd_fused_supported": 1}``)"""

    def __init__(self, returns=None):
        self.protos = parse_header()
        self.returns = dict(returns or {})
        # pcb_comm_init hands a handle back through its out-parameter: the double writes a non-null token there
        self.effects = {"pcb_comm_init": lambda uid, rank, world, out: setattr(out._obj, "value", 0x1000)}
        self.calls: List[str] = []
        self._fns: Dict[str, _Fn] = {}

    def __getattr__(self, name):
        if name.startswith("_") or name in ("protos", "returns", "calls", "effects"):
            raise AttributeError(name)
        if name not in self.protos:
            raise AttributeError(f"{name} is not declared in include/pcb200.h")
        if name not in self._fns:
            self._fns[name] = _Fn(self, name, *self.protos[name])
        return self._fns[name]


class _NoSide:
    """stand-in for the weight-gradient side stream of ``_mednext_bwd._Side`` (no CUDA streams on the CPU)"""
    enabled = False
    stream = None

    def __init__(self, device):
        pass

    def fork(self, *tensors):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def join(self):
        pass


def install(monkeypatch, returns=None) -> DryRunLib:
    from pytorch_connectomics_b200 import _lib as L
    from pytorch_connectomics_b200.architectures import _mednext_bwd as B
    lib = DryRunLib(returns)
    monkeypatch.setattr(L, "lib", lambda: lib)
    monkeypatch.setattr(L, "require_device", lambda t, what: None)
    monkeypatch.setattr(L, "stream_ptr", lambda device=None: ctypes.c_void_p(0))
    monkeypatch.setattr(B, "_Side", _NoSide)
    return lib

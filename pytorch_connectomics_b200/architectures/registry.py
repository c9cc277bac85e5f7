"""Architecture registry seam (drop-in for ``connectomics/models/architectures/registry.py:17-119``).

Same public functions, same semantics: duplicate registration warns (``UserWarning``) and
overwrites; unknown names raise ``ValueError`` listing what is available; ``unregister`` of an
unknown name raises ``ValueError``.  Backed by one table object so tests can snapshot/restore it.
"""

from __future__ import annotations

import warnings
from typing import Callable, Dict, List


class _Table:
    def __init__(self) -> None:
        self.builders: Dict[str, Callable] = {}

    def add(self, name: str, fn: Callable) -> Callable:
        if name in self.builders:
            warnings.warn(f"Architecture '{name}' already registered. Overwriting previous registration.",
                          UserWarning, stacklevel=3)
        self.builders[name] = fn
        return fn

    def get(self, name: str) -> Callable:
        try:
            return self.builders[name]
        except KeyError:
            raise ValueError(
                f"Architecture '{name}' not found.\nAvailable architectures: {sorted(self.builders)}\n"
                "Register new architectures with @register_architecture decorator.") from None

    def drop(self, name: str) -> None:
        if self.builders.pop(name, None) is None:
            raise ValueError(f"Architecture '{name}' not registered.")


_TABLE = _Table()


def register_architecture(name: str):
    return lambda builder_fn: _TABLE.add(name, builder_fn)


def get_architecture_builder(name: str) -> Callable:
    return _TABLE.get(name)


def list_architectures() -> List[str]:
    return sorted(_TABLE.builders)


def is_architecture_available(name: str) -> bool:
    return name in _TABLE.builders


def unregister_architecture(name: str) -> None:
    _TABLE.drop(name)


def get_architecture_info() -> Dict[str, Dict[str, str]]:
    return {name: {"name": name, "module": fn.__module__,
                   "doc": (fn.__doc__ or "").strip() or "No documentation"}
            for name, fn in _TABLE.builders.items()}


__all__ = ["register_architecture", "get_architecture_builder", "list_architectures",
           "is_architecture_available", "unregister_architecture", "get_architecture_info"]

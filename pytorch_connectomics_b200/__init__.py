"""pcb200 — B200-native engine for the PyTorch Connectomics hot path.

Drop-in replacements, behind the reference's own seams:
  * ``pytorch_connectomics_b200.architectures`` — architecture registry + MedNeXt builders
    (``connectomics.models.architectures``; ``connectomics.models.build``)
  * ``pytorch_connectomics_b200.inference`` — sliding-window engine
    (``connectomics.inference.window``)
Everything numeric runs in hand-written sm_100a CUDA behind the C ABI in ``include/pcb200.h``.
"""

__version__ = "0.1.0"

from . import _lib  # noqa: F401

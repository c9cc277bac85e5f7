"""Data-parallel training step plumbing (DDP seam, ``training/lightning/trainer.py:231-256``)."""
from .ddp import FlatGradArena, allreduce_gradients  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401

"""Data-parallel training step plumbing (DDP seam, ``training/lightning/trainer.py:231-256``)."""
from .ddp import FlatGradArena, allreduce_gradients, broadcast_parameters  # noqa: F401
from .data_parallel import (ArenaDataParallel, GradSegment, allreduce_mean_hook, allreduce_sum_hook,  # noqa: F401
                            bf16_compress_hook, make_arena_ddp_strategy, plan_segments)
from .deep_supervision import (deep_supervision_loss, deep_supervision_weights, match_target_to_output,  # noqa: F401
                               split_outputs)
from .graph import GraphedTrainStep  # noqa: F401
from .optim import FusedAdamW, build_fused_adamw, reference_param_groups  # noqa: F401
from .step import ArenaTrainStep  # noqa: F401

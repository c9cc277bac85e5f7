"""Sliding-window inference on the B200 engine (drop-in for ``connectomics.inference`` window API)."""

from .window import (EagerSlidingWindowEngine, apply_border_mask, build_sliding_accumulator_weight_maps,
                     build_sliding_importance_map, build_sliding_inferer, compute_importance_map,
                     compute_scan_interval, dense_patch_slices, is_distance_transform_blending,
                     normalize_weighted_accumulator, resolve_border_mask, resolve_inferer_overlap,
                     resolve_inferer_roi_size, resolve_model_output_dtype)

from .lazy import (ArrayVolumeAccessor, build_accessor, lazy_predict_region, lazy_predict_volume, lazy_sliding_window,
                   lazy_window_records, register_accessor_factory)
from .artifact import (PredictionArtifactMetadata, build_prediction_artifact_metadata, read_prediction_artifact,
                       write_prediction_artifact)
from .chunked import (ChunkRef, build_chunk_grid, chunks_for_rank, is_chunked_inference_enabled,
                      is_external_chunk_sharding_enabled, resolve_chunk_shape, resolve_external_chunk_shard,
                      resolve_halo_region, run_chunked_prediction, run_chunked_prediction_inference, stitch_chunks)
from . import chunk_grid, lazy_distributed, output  # noqa: F401  (reference module names: inference.chunk_grid / .output)

from .sharded import SlabPlan, ZSlabShardedEngine, exchange_overlaps, plan_z_slabs
from .tta import TTAEnsemble, apply_view, resolve_tta_augmentation_combinations

__all__ = ["TTAEnsemble", "apply_view", "resolve_tta_augmentation_combinations", "SlabPlan", "ZSlabShardedEngine", "exchange_overlaps", "plan_z_slabs", "lazy_predict_region", "lazy_predict_volume", "lazy_sliding_window", "lazy_window_records", "ChunkRef",
           "build_chunk_grid", "chunks_for_rank", "resolve_chunk_shape", "resolve_external_chunk_shard",
           "resolve_halo_region", "run_chunked_prediction", "stitch_chunks", "is_chunked_inference_enabled",
           "is_external_chunk_sharding_enabled", "run_chunked_prediction_inference",
           "EagerSlidingWindowEngine", "apply_border_mask", "build_sliding_accumulator_weight_maps",
           "build_sliding_importance_map", "build_sliding_inferer", "compute_importance_map",
           "compute_scan_interval", "dense_patch_slices", "is_distance_transform_blending",
           "normalize_weighted_accumulator", "resolve_border_mask", "resolve_inferer_overlap",
           "resolve_inferer_roi_size", "resolve_model_output_dtype"]

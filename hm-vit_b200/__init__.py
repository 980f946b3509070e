"""hmvit_b200 -- B200 (sm_100a) native HM-ViT multi-agent BEV fusion hot path.

The directory is called `hm-vit_b200` (not importable by name); load it with `hmvit_loader.load()`
from the repo root, which registers it as the module `hmvit_b200`.
"""
from . import _lib, ops, training  # noqa: F401
from .fusion import (HeteroAttention, HeteroFeedForward, HeteroFusion, HeteroFusionBlock,  # noqa: F401
                     HeteroLayerNorm, HeteroPreNormResidual, SpatialTransformation,
                     get_roi_and_cav_mask, regroup)
from .decoder import HeteroDecoder, NaiveDecoder  # noqa: F401
from .model import BevformerPointPillarHetero, combine_features, unpad_features, unpad_mode_encoding  # noqa: F401
from .postprocess import VoxelPostprocessor  # noqa: F401
from . import encoders  # noqa: F401
from .encoders import CvtCameraEncoder, PointPillar, build_config3_model, config3_args  # noqa: F401
from .build import build_extension  # noqa: F401
from .sharding import max_over_ranks, scene_shard  # noqa: F401
from .distributed import FlatGradAllReduce  # noqa: F401

__all__ = ["HeteroFusion", "HeteroFusionBlock", "HeteroAttention", "HeteroLayerNorm", "HeteroFeedForward",
           "HeteroPreNormResidual", "SpatialTransformation", "get_roi_and_cav_mask", "regroup",
           "HeteroDecoder", "NaiveDecoder", "VoxelPostprocessor", "BevformerPointPillarHetero", "combine_features", "unpad_features",
           "unpad_mode_encoding", "encoders", "PointPillar", "CvtCameraEncoder", "build_config3_model", "config3_args", "build_extension", "ops", "training", "FlatGradAllReduce"]

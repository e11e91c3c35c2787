"""shasta_b200 — B200-native (sm_100a) implementation of ShaSTA's affinity-estimation hot path.

Public surface mirrors the reference's det3d interface for this path:
    from shasta_b200 import build_track, build_simp_track, TRACK, SECOND_STAGE, Shasta, BEVFeatureExtractor
The compute lives in ``libshasta_b200.so`` (C ABI: include/shasta_b200.h), built in-tree by ``shasta_b200.build``.
"""
from .registry import (BACKBONES, NECKS, READERS, SECOND_STAGE, TRACK, Registry, build_backbone, build_from_cfg,  # noqa: F401
                       build_neck, build_reader, build_second_stage_module, build_simp_track, build_track)
from .bev_extractor import BEVFeatureExtractor  # noqa: F401
from .shasta import Shasta, load_matching_state_dict  # noqa: F401

__all__ = ["Registry", "build_from_cfg", "TRACK", "SECOND_STAGE", "READERS", "BACKBONES", "NECKS", "build_track",
           "build_simp_track", "build_second_stage_module", "build_reader", "build_backbone", "build_neck",
           "Shasta", "BEVFeatureExtractor", "load_matching_state_dict"]

"""revers_o_b200 — B200-native (sm_100a) implementation of the region-similarity hot path of
kolenyo2099/revers-o: mask-pooled region embeddings (K1), exact cosine top-k search (K2) and the
cross-GPU top-k merge (K3), behind the reference's `core_system.SimpleReverso` entry points.

Importing the package does not load the CUDA library; the first compute call does, and fails loudly
if it is missing (there is no CPU fallback).  Build: `python -m revers_o_b200.build`.
"""
from ._lib import RVO_MAX_K, RVO_SMALL_Q, RvoError  # noqa: F401

__all__ = ["RvoError", "RVO_MAX_K", "RVO_SMALL_Q"]

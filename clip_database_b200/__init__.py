"""clip_database_b200 — B200-native brute-force KNN path for CLIP-database.

One hot path of droon/CLIP-database, rebuilt for sm_100a behind a C ABI
(``include/clipdb.h``): the exhaustive cosine scan + top-k that
``ImageDatabase.search()`` runs through sqlite-vec (image_database.py:1564-1583),
plus the query blend / negative-prompt arithmetic before it (:1378-1398, :545-604).

  GpuIndex        one GPU's resident row store + kernels (ctypes over the C ABI)
  ImageDatabase   the reference's ``search()`` surface on top of it
  ShardedIndex    row-sharded multi-GPU search, one process per GPU (torch.distributed)
  MultiGpuIndex   the same sharding from ONE process (clip_database_b200.multigpu)
  dropin.install  patch the reference class's own search() to use this path

Importing this package does not need a GPU; creating a ``GpuIndex`` does, and
fails loudly without one (no CPU fallback).
"""
from .index import GpuIndex, SearchResult  # noqa: F401
from .database import Embedder, ImageDatabase, filter_duplicates, like_prefix_mask  # noqa: F401

__all__ = ["GpuIndex", "SearchResult", "ImageDatabase", "Embedder", "filter_duplicates",
           "like_prefix_mask"]

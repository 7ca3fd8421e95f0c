"""Drop the B200 search path into the reference class itself.

``install(image_database.ImageDatabase)`` replaces ONLY the reference's ``search()``
(image_database.py:1308-1658); model loading, scanning, the interactive REPL
(:2070-2299, which calls ``self.search(...)``), the HTML gallery and the CLI keep running the
reference's own code, now answered from HBM.  The embeddings still come from the reference's
``_get_text_embedding`` / ``_get_image_embedding`` (:509-543, :443-463).

    import image_database, clip_database_b200.dropin as dropin
    dropin.install(image_database.ImageDatabase, device=0)           # once
    db = image_database.ImageDatabase("images.db")                    # the reference, unchanged
    db.search("a red car", k=20, negative_query="people")              # scan + top-k on the GPU

The resident store is built on the first search of each instance.  Before every later search
``ImageDatabase.refresh()`` checks SQLite's change counter (O(1)) and, when something was committed,
reconciles the store with the mapping tables: appended rows are streamed in, rows orphaned by a
re-scan of a modified file (``INSERT OR REPLACE INTO images`` re-keys it) are retired, re-embedded
rows are re-read — what the reference sees by reopening the database for every search.
``uninstall`` restores the original method.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

from .database import Embedder, ImageDatabase


class _ReferenceEmbedder(Embedder):
    """The reference instance's own model calls behind the ``Embedder`` interface."""

    def __init__(self, ref):
        self.ref = ref

    def text(self, query: str):
        return self.ref._get_text_embedding(query)

    def image(self, path: str):
        return self.ref._get_image_embedding(path)


def install(reference_cls, device: int = 0, **database_kwargs):
    """Patch ``reference_cls.search`` in place; returns the class.  ``database_kwargs`` go to
    ``clip_database_b200.ImageDatabase`` (``nan_policy``, ``binary_score_mode``, ``batch_store``,
    ``devices=[...]`` to row-shard the store over several GPUs)."""
    if getattr(reference_cls, "_b200_original_search", None) is not None:
        uninstall(reference_cls)
    original = reference_cls.search

    def search(self, query: str, k: int = 10, is_image_path: bool = False,
               query2: str = None, is_image_path2: bool = False,
               weights: Tuple[float, float] = (0.5, 0.5),
               negative_query: str = None, negative_is_image: bool = False,
               negative_weight: float = 0.5,
               negative_queries: List[str] = None, negative_is_images: List[bool] = None,
               negative_weights: List[float] = None,
               filter_folders: List[str] = None,
               profile: bool = False,
               show_duplicates: bool = False) -> List[Tuple[str, float]]:
        db: Optional[ImageDatabase] = getattr(self, "_b200", None)
        if db is None or db.db_path != self.db_path:
            if db is not None:
                db.close()
            db = ImageDatabase(self.db_path, device=device, embedder=_ReferenceEmbedder(self), **database_kwargs)
            self._b200 = db
        else:
            db.refresh()           # the reference opens a fresh connection per search and sees new rows
        return db.search(query, k=k, is_image_path=is_image_path, query2=query2, is_image_path2=is_image_path2,
                         weights=weights, negative_query=negative_query, negative_is_image=negative_is_image,
                         negative_weight=negative_weight, negative_queries=negative_queries,
                         negative_is_images=negative_is_images, negative_weights=negative_weights,
                         filter_folders=filter_folders, profile=profile, show_duplicates=show_duplicates)

    search.__doc__ = original.__doc__
    reference_cls._b200_original_search = original
    reference_cls.search = search
    return reference_cls


def uninstall(reference_cls):
    original = getattr(reference_cls, "_b200_original_search", None)
    if original is not None:
        reference_cls.search = original
        reference_cls._b200_original_search = None
    return reference_cls

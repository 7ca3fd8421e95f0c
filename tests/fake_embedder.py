"""A deterministic stand-in for the SigLIP model (tests of the --embedder hook): the embedding of a string is
a seeded unit vector derived from its bytes."""
import zlib

import numpy as np


class HashEmbedder:
    dim = 1152

    def _vec(self, s: str) -> np.ndarray:
        v = np.random.default_rng(zlib.crc32(s.encode("utf-8"))).standard_normal(self.dim, dtype=np.float32)
        return v / np.linalg.norm(v)

    def text(self, query: str):
        return self._vec("text:" + query)

    def image(self, path: str):
        return self._vec("image:" + path)


class NotAnEmbedder:
    def text(self, query):
        return None

"""Shared test fixtures: a tiny text corpus and a deterministic hash embedder
(stands in for the reference's LocalNLPModels: ``embed_single`` / ``embed``)."""

from __future__ import annotations

import zlib
from typing import List

import numpy as np

_TOPICS = {
    "python": "python programming language interpreter dynamic typing scripts",
    "java": "java programming language virtual machine enterprise static typing",
    "snake": "python snake reptile animal jungle constrictor",
    "gpu": "gpu kernel cuda tensor memory bandwidth warp shared memory",
    "search": "search retrieval index ranking bm25 sparse dense vector",
    "quant": "binary quantization hamming distance int8 rescoring embedding vector",
    "db": "database redis postgres storage index query transaction",
    "cook": "recipe cooking pasta tomato garlic olive oil kitchen",
}
_EXTRA = ["fast", "slow", "modern", "classic", "large", "small", "simple", "robust"]

CORPUS_TEXTS: List[str] = []
_keys = sorted(_TOPICS)
for _i in range(48):
    a = _TOPICS[_keys[_i % len(_keys)]].split()
    b = _TOPICS[_keys[(_i * 3 + 1) % len(_keys)]].split()
    words = a[: 3 + _i % 4] + b[: 1 + _i % 3] + [_EXTRA[_i % len(_EXTRA)], _EXTRA[(_i * 5 + 2) % len(_EXTRA)]]
    if _i % 5 == 0:
        words += a[:2]  # repeated terms -> tf > 1
    CORPUS_TEXTS.append(" ".join(words).capitalize() + f". Note {_i}!")

QUERY_TEXTS: List[str] = [
    "python programming",
    "snake animal in the jungle",
    "gpu kernel memory bandwidth",
    "binary quantization hamming rescoring",
    "redis database index query",
    "fast simple search ranking",
    "pasta recipe with garlic",
    "java virtual machine typing",
    "unknownterm zzz",
    "vector embedding retrieval dense sparse",
]


class HashEmbedder:
    """token -> fixed pseudo-random vector; text -> L2-normalised sum (float32)."""

    def __init__(self, dim: int = 64, seed: int = 77) -> None:
        self.dim = dim
        self.seed = seed

    def _tok(self, tok: str) -> np.ndarray:
        rng = np.random.default_rng([self.seed, zlib.crc32(tok.encode("utf-8"))])
        return rng.standard_normal(self.dim).astype(np.float32)

    def embed_single(self, text: str) -> List[float]:
        toks = [t for t in "".join(c if c.isalnum() else " " for c in text.lower()).split() if len(t) > 1]
        v = np.zeros(self.dim, dtype=np.float32)
        for t in toks:
            v += self._tok(t)
        n = np.linalg.norm(v)
        if n == 0:
            v[0] = 1.0
            n = 1.0
        return (v / n).astype(np.float32).tolist()

    def embed(self, texts: List[str]) -> List[List[float]]:
        return [self.embed_single(t) for t in texts]

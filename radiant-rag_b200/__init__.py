"""radiant-rag_b200: B200-native retrieval hot path for Radiant RAG.

Binary-quantised dense search (exact Hamming top-k -> int8 / float32 rescoring), BM25
sparse scoring and Reciprocal Rank Fusion behind the reference's own interfaces
(``BaseVectorStore`` / ``PersistentBM25Index`` / ``DenseRetrievalAgent`` /
``BM25RetrievalAgent`` / ``RRFAgent``).  Python host code (torch tensors for device
memory, streams and torch.distributed) calls hand-written sm_100a CUDA kernels
through the C ABI of ``librr_b200.so`` (include/radiant_rag_b200.h).

The directory name carries a hyphen, so it is imported as ``radiant_rag_b200`` (a
two-line shim package next to it extends its ``__path__`` here).
"""

from . import _lib  # noqa: F401
from ._lib import RadiantB200Error, LIB_PATH  # noqa: F401
from .base import BaseVectorStore, StoredDoc  # noqa: F401
from .config import BM25Config, QuantizationConfig, RetrievalConfig  # noqa: F401

__all__ = [
    "RadiantB200Error", "LIB_PATH", "BaseVectorStore", "StoredDoc",
    "BM25Config", "QuantizationConfig", "RetrievalConfig",
    "B200VectorStore", "DenseIndex", "BM25Index", "PersistentBM25Index", "Bm25DeviceIndex",
    "DenseRetrievalAgent", "BM25RetrievalAgent", "RRFAgent",
]


def __getattr__(name):  # lazy: these import torch
    if name in ("B200VectorStore",):
        from .vector_store import B200VectorStore
        return B200VectorStore
    if name in ("DenseIndex",):
        from .index import DenseIndex
        return DenseIndex
    if name in ("BM25Index", "PersistentBM25Index", "Bm25DeviceIndex"):
        from . import bm25_index
        return getattr(bm25_index, name)
    if name in ("DenseRetrievalAgent", "BM25RetrievalAgent", "RRFAgent", "AgentResult"):
        from . import agents
        return getattr(agents, name)
    raise AttributeError(name)

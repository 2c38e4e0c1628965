"""Host-side mirror of the reference storage interface (the drop-in boundary).

Same names, argument meaning and error behaviour as reference
radiant/storage/base.py: ``StoredDoc`` :23-37 (identity is ``doc_id``),
``BaseVectorStore`` :40-325 (15 abstract methods + ``retrieve_by_embedding_quantized``),
``_default_make_doc_id`` :311-325 (sha256 of content + sorted-key JSON of meta).

When the reference package is importable the classes here are registered as
virtual subclasses of the reference ABCs, so ``isinstance`` checks in a host
application keep working; nothing here imports the reference at module import.
"""

from __future__ import annotations

import hashlib
import json
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple


@dataclass
class StoredDoc:
    """A stored document; equality and hash are by ``doc_id`` only."""

    doc_id: str
    content: str
    meta: Dict[str, Any]

    def __hash__(self) -> int:
        return hash(self.doc_id)

    def __eq__(self, other: object) -> bool:
        return hasattr(other, "doc_id") and hasattr(other, "content") and self.doc_id == other.doc_id


def normalize_doc_level(doc_level_filter: Optional[str]) -> Optional[str]:
    """Filter normalisation of reference radiant/storage/redis_store.py:669-677:
    child/leaves/leaf -> "child"; parent/parents -> "parent"; None/"all"/other -> None."""
    if not doc_level_filter:
        return None
    low = doc_level_filter.lower()
    if low in ("child", "leaves", "leaf"):
        return "child"
    if low in ("parent", "parents"):
        return "parent"
    return None


class BaseVectorStore(ABC):
    """Interface every storage backend implements (reference base.py:40-325)."""

    @abstractmethod
    def ping(self) -> bool: ...

    @abstractmethod
    def make_doc_id(self, content: str, meta: Optional[Dict[str, Any]] = None) -> str: ...

    @abstractmethod
    def upsert(self, doc_id: str, content: str, embedding: List[float],
               meta: Optional[Dict[str, Any]] = None) -> None: ...

    @abstractmethod
    def upsert_doc_only(self, doc_id: str, content: str, meta: Optional[Dict[str, Any]] = None) -> None: ...

    @abstractmethod
    def upsert_batch(self, documents: List[Dict[str, Any]]) -> int: ...

    @abstractmethod
    def upsert_doc_only_batch(self, documents: List[Dict[str, Any]]) -> int: ...

    @abstractmethod
    def get_doc(self, doc_id: str) -> Optional[StoredDoc]: ...

    @abstractmethod
    def has_embedding(self, doc_id: str) -> bool: ...

    @abstractmethod
    def delete_doc(self, doc_id: str) -> bool: ...

    @abstractmethod
    def retrieve_by_embedding(
        self,
        query_embedding: List[float],
        top_k: int,
        min_similarity: float = 0.0,
        ef_runtime: Optional[int] = None,
        language_filter: Optional[str] = None,
        doc_level_filter: Optional[str] = None,
    ) -> List[Tuple[StoredDoc, float]]: ...

    def retrieve_by_embedding_quantized(
        self,
        query_embedding: List[float],
        top_k: int,
        min_similarity: float = 0.0,
        rescore_multiplier: Optional[float] = None,
        use_rescoring: Optional[bool] = None,
        language_filter: Optional[str] = None,
        doc_level_filter: Optional[str] = None,
    ) -> List[Tuple[StoredDoc, float]]:
        """Default: fall back to the float path (reference base.py:242-249)."""
        return self.retrieve_by_embedding(
            query_embedding=query_embedding,
            top_k=top_k,
            min_similarity=min_similarity,
            language_filter=language_filter,
            doc_level_filter=doc_level_filter,
        )

    @abstractmethod
    def list_doc_ids(self, pattern: str = "*", limit: int = 10_000) -> List[str]: ...

    @abstractmethod
    def list_doc_ids_with_embeddings(self, limit: int = 10_000) -> List[str]: ...

    @abstractmethod
    def get_index_info(self) -> Dict[str, Any]: ...

    @abstractmethod
    def drop_index(self, delete_documents: bool = False) -> bool: ...

    @abstractmethod
    def count_documents(self) -> int: ...

    def _default_make_doc_id(self, content: str, meta: Optional[Dict[str, Any]] = None) -> str:
        meta_part = json.dumps(meta or {}, sort_keys=True, ensure_ascii=False)
        return hashlib.sha256((content + "\n" + meta_part).encode("utf-8")).hexdigest()


def register_with_reference() -> bool:
    """Make these classes virtual subclasses of the reference ABCs when the reference
    package is importable (no-op otherwise)."""
    try:
        from radiant.storage.base import BaseVectorStore as RefStore  # type: ignore
    except Exception:
        return False
    RefStore.register(BaseVectorStore)
    return True

"""The hot path's knobs, mirrored from the reference's frozen dataclasses
(reference radiant/config.py): ``QuantizationConfig`` :275-295, ``BM25Config`` :386-395,
``RetrievalConfig`` :419-431.  Same field names and defaults, so either these or the
reference's own instances can be handed to the stores/agents in this package (only
attribute access is used).  Loading them from YAML stays in the reference.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional


@dataclass(frozen=True)
class QuantizationConfig:
    enabled: bool = False
    precision: str = "both"  # "binary", "int8" or "both"
    rescore_multiplier: float = 4.0
    use_rescoring: bool = True
    int8_ranges_file: Optional[str] = None
    int8_on_disk_only: bool = True

    def __post_init__(self) -> None:
        # validation of reference radiant/storage/quantization.py:64-71
        if self.precision not in ("binary", "int8", "both"):
            raise ValueError(f"Invalid precision '{self.precision}'. Must be 'binary', 'int8', or 'both'")
        if self.rescore_multiplier < 1.0:
            raise ValueError(f"rescore_multiplier must be >= 1.0, got {self.rescore_multiplier}")


@dataclass(frozen=True)
class BM25Config:
    index_path: str = "./data/bm25_index"
    max_documents: int = 100_000
    auto_save_threshold: int = 100
    k1: float = 1.5
    b: float = 0.75


@dataclass(frozen=True)
class RetrievalConfig:
    dense_top_k: int = 10
    bm25_top_k: int = 10
    fused_top_k: int = 15
    rrf_k: int = 60
    min_similarity: float = 0.0
    search_scope: str = "leaves"


def quantization_from_yaml(source, backend: Optional[str] = None) -> QuantizationConfig:
    """The ``quantization:`` section the reference's ``load_config`` never reads
    (radiant/config.py:1179-1228 build Redis/Chroma/PgVector configs without it), as
    laid out in the reference's ``config_quantization_example.yaml``: one section per backend
    (``redis:`` / ``chroma:`` / ``pgvector:``), selected by ``storage.backend`` unless `backend` is
    given.  `source`: a path to the YAML file or the already-parsed mapping.  Unknown keys are
    ignored, missing ones keep the reference's defaults; ``RADIANT_QUANTIZATION_<KEY>`` environment
    variables override, in the reference's convention (config.py: RADIANT_<SECTION>_<KEY>)."""
    import os

    if isinstance(source, (str, os.PathLike)):
        import yaml

        with open(source, "r", encoding="utf-8") as f:
            data = yaml.safe_load(f) or {}
    else:
        data = dict(source or {})
    name = backend or (data.get("storage") or {}).get("backend") or "redis"
    section = ((data.get(name) or {}).get("quantization")) or data.get("quantization") or {}
    fields = {"enabled": bool, "precision": str, "rescore_multiplier": float, "use_rescoring": bool,
              "int8_ranges_file": str, "int8_on_disk_only": bool}
    kwargs = {}
    for key, typ in fields.items():
        val = section.get(key)
        env = os.environ.get(f"RADIANT_QUANTIZATION_{key.upper()}")
        if env is not None:
            val = env
        if val is None:
            continue
        if typ is bool and isinstance(val, str):
            val = val.strip().lower() in ("1", "true", "yes", "on")
        kwargs[key] = typ(val)
    return QuantizationConfig(**kwargs)

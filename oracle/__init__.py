"""CPU oracle for the Radiant RAG retrieval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``radiant-rag_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker or the
timed CPU arm, never as the product path.

Each function is a NumPy restatement of one row of SURVEY.md section 8(a) and
cites the reference file:line it follows (paths relative to the upstream repo
dshipley71/radiant-rag).

Parity status
-------------
* BM25 (R6/R7/R8), RRF (R10), rescoring (R3), the two-stage control flow and
  the exact cosine scan (R5) are PINNED: ``tests/golden/`` holds outputs of the
  reference's own Python code (``oracle/gen_golden.py`` imports it from
  ``/root/reference`` in the build container) and ``tests/test_oracle_golden.py``
  checks this oracle against them.
* ``ubinary`` / ``int8`` quantisation (R1/R2) and the Hamming top-k (R4) are
  **parity unpinned**: the arithmetic lives in the third-party dependency
  ``sentence-transformers`` (``sentence_transformers.quantization``; constraint
  ``>=3.2.0`` in requirements.txt:47, no lock file, not vendored, not installed
  here) and no reference code computes a Hamming distance at all
  (SURVEY.md section 0.2).  They are restated from the published algorithm and
  anchored on the reference's call sites and on the only checks the reference
  holds (tools/validate_quantization.py:142,159-160,169-170).
"""

from .quantize import (  # noqa: F401
    quantize_ubinary,
    quantize_int8,
    calculate_int8_ranges,
    quantize_int8_symmetric_query,
)
from .hamming import hamming_distances, hamming_topk  # noqa: F401
from .rescore import (  # noqa: F401
    rescore_f32, rescore_i8_exact, exact_cosine_topk, int8_exact_topk, int8_exact_topk_blas,
)
from .bm25 import tokenize, BM25Oracle  # noqa: F401
from .rrf import rrf_fuse  # noqa: F401
from .flow import two_stage_search  # noqa: F401

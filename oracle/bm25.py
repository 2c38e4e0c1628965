"""Oracle (test infrastructure): BM25 scoring, rows R6/R7/R8 of SURVEY.md 8(a).

Follows reference radiant/storage/bm25_index.py:
  * ``_tokenize`` :50-58 (lower-case, non-alnum -> space, split, drop len <= 1)
  * ``BM25Index._rebuild_index`` :100-137 (avgdl, df, idf = log((n-df+0.5)/(df+0.5)+1))
  * ``BM25Index.add_document`` :139-180 (incremental idf refresh of the new doc's
    own terms only - the stale-idf quirk; reproduced by ``from_reference`` which
    copies idf/avgdl from a reference object instead of recomputing)
  * ``BM25Index.search`` :218-270 (float64 accumulate in query-token order, repeats
    included; top-k; drop score <= 0)

The reference walks every document with ``list.count`` per query term; this
restatement walks an inverted CSR instead but performs the SAME IEEE-754 double
operations in the SAME order per (query token, document), so scores are
bit-identical (checked against the reference in tests/golden/bm25_*.json).

Canonical order: the reference's ``argpartition``/unstable ``argsort`` leaves ties
unspecified; the oracle (and the CUDA path) use (score desc, row asc).
"""

from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def tokenize(text: str) -> List[str]:
    """Restates reference bm25_index.py:50-58."""
    cleaned = []
    for ch in text.lower():
        cleaned.append(ch if ch.isalnum() else " ")
    return [tok for tok in "".join(cleaned).split() if len(tok) > 1]


class BM25Oracle:
    """Inverted-index BM25 with the reference's exact float64 arithmetic.

    Documents are rows 0..N-1 (insertion order of ``BM25Index.doc_ids``); terms
    are integer ids 0..V-1.
    """

    def __init__(
        self,
        doc_ptr: np.ndarray,
        doc_terms: np.ndarray,
        n_terms: int,
        k1: float = 1.5,
        b: float = 0.75,
        idf: Optional[np.ndarray] = None,
        avgdl: Optional[float] = None,
        only_terms: Optional[Iterable[int]] = None,
    ) -> None:
        """only_terms: index just these term ids (document lengths, avgdl and the document
        count still come from ALL tokens, and df / idf of an indexed term only depend on its own
        postings, so scores of queries made of these terms are unchanged) - lets a full-size
        corpus be checked on a sample of queries without sorting every token."""
        self.k1 = float(k1)
        self.b = float(b)
        doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
        doc_terms = np.asarray(doc_terms, dtype=np.int64)
        self.n_docs = int(doc_ptr.size - 1)
        self.n_terms = int(n_terms)
        self.doc_len = np.diff(doc_ptr).astype(np.int64)
        # bm25_index.py:117: avgdl = sum(len) / n  (Python ints -> one double divide)
        if avgdl is None:
            avgdl = (int(self.doc_len.sum()) / self.n_docs) if self.n_docs else 0.0
        self.avgdl = float(avgdl)

        # inverted CSR: sort (term, row) pairs, collapse repeats into tf
        rows = np.repeat(np.arange(self.n_docs, dtype=np.int64), self.doc_len)
        if only_terms is not None:
            keep = np.isin(doc_terms, np.fromiter((int(t) for t in only_terms), dtype=np.int64))
            rows, doc_terms = rows[keep], doc_terms[keep]
        key = doc_terms * np.int64(max(self.n_docs, 1)) + rows
        del rows
        key.sort(kind="stable")
        if key.size:
            head = np.empty(key.size, dtype=bool)
            head[0] = True
            np.not_equal(key[1:], key[:-1], out=head[1:])
            first = np.flatnonzero(head)
            uniq = key[first]
            tf = np.diff(np.append(first, key.size))
        else:
            uniq, tf = key, np.zeros(0, dtype=np.int64)
        del key
        self.post_term = (uniq // max(self.n_docs, 1)).astype(np.int64)
        self.post_row = (uniq % max(self.n_docs, 1)).astype(np.int64)
        self.post_tf = tf.astype(np.int64)
        self.df = np.bincount(self.post_term, minlength=self.n_terms).astype(np.int64)
        self.term_ptr = np.zeros(self.n_terms + 1, dtype=np.int64)
        np.cumsum(self.df, out=self.term_ptr[1:])

        if idf is None:
            # bm25_index.py:131-135, scalar np.log per term exactly as the reference
            idf = np.zeros(self.n_terms, dtype=np.float64)
            n = self.n_docs
            for t in np.nonzero(self.df)[0]:
                d = int(self.df[t])
                idf[t] = np.log((n - d + 0.5) / (d + 0.5) + 1.0)
        self.idf = np.asarray(idf, dtype=np.float64)
        # terms with df == 0 are "not in self.idf" (bm25_index.py:238-239)
        self.known = self.df > 0

    # ---- construction helpers -------------------------------------------------
    @classmethod
    def from_token_lists(
        cls, docs: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75
    ) -> Tuple["BM25Oracle", Dict[str, int]]:
        vocab: Dict[str, int] = {}
        ptr = [0]
        terms: List[int] = []
        for toks in docs:
            for t in toks:
                terms.append(vocab.setdefault(t, len(vocab)))
            ptr.append(len(terms))
        return cls(np.asarray(ptr), np.asarray(terms, dtype=np.int64), len(vocab), k1, b), vocab

    @classmethod
    def from_reference(cls, ref_index) -> Tuple["BM25Oracle", Dict[str, int]]:
        """Mirror a reference ``BM25Index`` INCLUDING its current (possibly stale)
        idf table and incrementally-updated avgdl (bm25_index.py:162-177)."""
        inst, vocab = cls.from_token_lists(ref_index.doc_tokens, ref_index.k1, ref_index.b)
        idf = np.zeros(len(vocab), dtype=np.float64)
        known = np.zeros(len(vocab), dtype=bool)
        for term, tid in vocab.items():
            if term in ref_index.idf:
                idf[tid] = float(ref_index.idf[term])
                known[tid] = True
        inst.idf = idf
        inst.known = known
        inst.avgdl = float(ref_index.avgdl)
        return inst, vocab

    # ---- scoring --------------------------------------------------------------
    def impacts(self, t: int) -> Tuple[np.ndarray, np.ndarray]:
        """(rows, idf_t * num/den) for term t, in the reference's op order
        (bm25_index.py:252-255)."""
        lo, hi = self.term_ptr[t], self.term_ptr[t + 1]
        rows = self.post_row[lo:hi]
        tf = self.post_tf[lo:hi].astype(np.float64)
        dl = self.doc_len[rows].astype(np.float64)
        num = tf * (self.k1 + 1)
        den = tf + self.k1 * ((1 - self.b) + (self.b * dl) / self.avgdl)
        return rows, self.idf[t] * (num / den)

    def scores(self, query_terms: Iterable[int]) -> np.ndarray:
        s = np.zeros(self.n_docs, dtype=np.float64)
        for t in query_terms:
            if t < 0 or t >= self.n_terms or not self.known[t]:
                continue
            rows, imp = self.impacts(int(t))
            s[rows] += imp  # rows are unique within a term
        return s

    def search(self, query_terms: Iterable[int], top_k: int) -> Tuple[np.ndarray, np.ndarray]:
        """-> (rows int64 [M], scores f64 [M]), M <= top_k, (score desc, row asc), score > 0."""
        s = self.scores(query_terms)
        pos = np.nonzero(s > 0)[0]
        if pos.size == 0:
            return np.empty(0, np.int64), np.empty(0, np.float64)
        if pos.size > top_k:
            # everything at or above the k-th largest score, then the canonical order among those
            # (same result as sorting all positive scores; the reference also partitions first,
            # bm25_index.py:258-262)
            sp = s[pos]
            kth = np.partition(sp, pos.size - top_k)[pos.size - top_k]
            pos = pos[sp >= kth]
        order = pos[np.lexsort((pos, -s[pos]))][:top_k]
        return order.astype(np.int64), s[order]

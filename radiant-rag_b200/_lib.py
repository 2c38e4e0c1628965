"""ctypes binding of librr_b200.so (the C ABI declared in include/radiant_rag_b200.h).

There is NO CPU fallback: if the library is missing or a call fails this module
raises, and every product path in this package goes through it.
ctypes releases the GIL for the duration of each foreign call, so the dense and
BM25 agents can be driven from two Python threads as the reference's orchestrator
does (radiant/orchestrator.py:994-998).
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path
from typing import Optional

PKG_DIR = Path(__file__).resolve().parent
# RR_B200_LIB points at an alternative build of the same library (kernel-variant experiments)
LIB_PATH = Path(os.environ["RR_B200_LIB"]) if os.environ.get("RR_B200_LIB") else PKG_DIR / "librr_b200.so"

RR_F32 = 0
RR_I8 = 1
RR_MAX_K = 1024
RR_MAX_WORDS = 32

_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_u8 = C.c_uint8
_u64 = C.c_uint64
_f64 = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/radiant_rag_b200.h
PROTOTYPES = {
    "rr_abi_version": (_i32, []),
    "rr_last_error": (C.c_char_p, []),
    "rr_init": (_i32, [_i32]),
    "rr_sm_count": (_i32, []),
    "rr_quantize_ubinary": (_i32, [_p, _i64, _i32, _p, _i32, _p]),
    "rr_quantize_int8": (_i32, [_p, _i64, _i32, _p, _p, _p]),
    "rr_hamming_topk_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "rr_hamming_topk": (_i32, [_p, _i64, _i32, _p, _u8, _u8, _p, _i32, _i32, _i64, _p, _p, _p, _sz, _p]),
    "rr_unpack_codes_pm1": (_i32, [_p, _i64, _i32, _i32, _p, _p]),
    "rr_tc_search_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "rr_hamming_topk_tc": (_i32, [_p, _i64, _i32, _p, _u8, _u8, _p, _i32, _i32, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "rr_int8_search_topk_tc": (_i32, [_p, _i64, _i32, _p, _u8, _u8, _p, _i32, _i32, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "rr_tc_dense_keys": (_i32, [_p, _i64, _i32, _p, _i32, _p, _p]),
    "rr_tc_timing": (_i32, [_i32]),
    "rr_tc_last_timing_ms": (_i32, [_p]),
    "rr_rescore_f32": (_i32, [_p, _i32, _i32, _p, _i32, _i64, _i64, _p, _i32, _i32, _f64, _p, _p, _p, _p]),
    "rr_score_candidates_f32": (_i32, [_p, _i32, _i32, _p, _i32, _i64, _i64, _p, _i32, _p, _p]),
    "rr_rank_scored_f32": (_i32, [_p, _p, _i32, _i32, _i32, _f64, _p, _p, _p, _p]),
    "rr_rescore_i8": (_i32, [_p, _i32, _i32, _p, _i64, _i64, _p, _i32, _i32, _p, _p, _p, _p]),
    "rr_exact_search_f32_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "rr_exact_search_f32": (_i32, [_p, _i64, _i32, _p, _u8, _u8, _p, _i32, _i32, _f64, _i64, _p, _p, _p, _p, _sz, _p]),
    "rr_row_inv_norms_f32": (_i32, [_p, _i64, _i32, _p, _p]),
    "rr_exact_search_f32_tc_supported": (_i32, [_i64, _i32, _i32, _i32]),
    "rr_exact_search_f32_tc_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "rr_exact_search_f32_tc": (_i32, [_p, _p, _i64, _i32, _p, _u8, _u8, _p, _i32, _i32, _f64, _i64, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "rr_int8_search_topk_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "rr_int8_search_topk": (_i32, [_p, _i64, _i32, _p, _u8, _u8, _p, _i32, _i32, _i64, _p, _p, _p, _sz, _p]),
    "rr_bm25_topk_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "rr_bm25_topk": (_i32, [_p, _p, _p, _i32, _i32, _i32, _i64, _p, _i32, _i32, _i32, _i64, _p, _p, _p, _p, _sz, _p]),
    "rr_bm25_fast_workspace_bytes": (_sz, [_i32, _i32, _i64, _i32, _i32]),
    "rr_bm25_fast_max_head": (_i32, [_i32]),
    "rr_bm25_topk_fast": (_i32, [_p, _p, _p, _p, _i32, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i64, _p, _i32, _i32, _i32, _i64,
                                 _p, _p, _p, _p, _p, _p, _sz, _p]),
    "rr_bm25_timing": (_i32, [_i32]),
    "rr_bm25_last_timing_ms": (_i32, [_p]),
    "rr_bm25_impacts": (_i32, [_p, _p, _p, _i64, _f64, _f64, _f64, _p, _p]),
    "rr_rrf_fuse": (_i32, [_p, C.POINTER(_i32), _i32, _i32, _f64, _i32, _p, _p, _p, _p]),
    "rr_rrf_fuse_runs": (_i32, [C.POINTER(_p), C.POINTER(_i32), _i32, _i32, _f64, _i32, _p, _p, _p, _p]),
    "rr_merge_hamming": (_i32, [_p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "rr_pack_hamming": (_i32, [_p, _p, _i64, _p, _p]),
    "rr_merge_hamming_gathered": (_i32, [_p, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "rr_merge_scores_f64_gathered": (_i32, [_p, _i32, _i32, _i32, _i32, _p, _p, _p, _p]),
    "rr_merge_scores_f64": (_i32, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _p]),
    "rr_merge_scores_i32": (_i32, [_p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "rr_probe_popc": (_i32, [_i32, C.POINTER(_f64), _p]),
    "rr_probe_gather": (_i32, [_p, _i32, _i64, _i32, _p, _i32, _i32, _p, C.POINTER(_f64), _p]),
    "rr_probe_smem": (_i32, [_i32, C.POINTER(_f64), _p]),
    "rr_probe_i8_mma": (_i32, [_i32, _i32, C.POINTER(_f64), _p]),
    "rr_synth_rows_f32": (_i32, [_p, _i64, _i64, _i32, _u64, _i32, _p]),
    "rr_synth_query_rows_f32": (_i32, [_p, _i64, _i64, _i32, _u64, _i64, _i32, _p]),
    "rr_synth_doc_lengths": (_i32, [_p, _i64, _i64, _u64, _i32, _p]),
    "rr_synth_zipf_tokens": (_i32, [_p, _i64, _i64, _u64, _p, _i32, _p]),
}

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()
_inited_devices = set()

# Kernel-launch counter: bench.py reports how many of OUR kernels ran in the timed region.
launch_count = 0
_KERNELS_PER_CALL = {
    "rr_quantize_ubinary": 1, "rr_quantize_int8": 1, "rr_hamming_topk": 2, "rr_rescore_f32": 1,
    "rr_score_candidates_f32": 1, "rr_rank_scored_f32": 1, "rr_rescore_i8": 1,
    "rr_exact_search_f32": 2, "rr_exact_search_f32_tc": 6, "rr_row_inv_norms_f32": 1, "rr_int8_search_topk": 2, "rr_bm25_topk": 2, "rr_bm25_topk_fast": 4, "rr_bm25_impacts": 1,
    "rr_rrf_fuse": 1, "rr_rrf_fuse_runs": 1, "rr_merge_hamming": 1, "rr_merge_scores_f64": 1, "rr_merge_scores_f64_gathered": 1, "rr_merge_scores_i32": 1,
    "rr_pack_hamming": 1, "rr_merge_hamming_gathered": 1,
    "rr_unpack_codes_pm1": 1, "rr_tc_dense_keys": 1, "rr_hamming_topk_tc": 5, "rr_int8_search_topk_tc": 5,
    "rr_synth_rows_f32": 1, "rr_synth_query_rows_f32": 1, "rr_synth_doc_lengths": 1,
    "rr_synth_zipf_tokens": 1,
}


class RadiantB200Error(RuntimeError):
    """Raised when the CUDA extension is missing or a C-ABI call fails."""


def load() -> C.CDLL:
    """Load librr_b200.so (works without a GPU; kernels need rr_init on a B200)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise RadiantB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback."
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().rr_last_error().decode("utf-8", "replace")


def init(device: int = 0) -> None:
    """rr_init for `device` once per process."""
    if device in _inited_devices:
        return
    lib = load()
    rc = lib.rr_init(int(device))
    if rc != 0:
        raise RadiantB200Error(f"rr_init({device}) failed ({rc}): {last_error()}")
    _inited_devices.add(device)


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point; raise RuntimeError with the C error string
    on failure (the reference's agents turn exceptions into an empty result,
    radiant/agents/base_agent.py:548-576)."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RadiantB200Error(f"{name} failed ({rc}): {last_error()}")
    launch_count += _KERNELS_PER_CALL.get(name, 0)


def ptr(t) -> int:
    """Device (or host) address of a torch tensor / None."""
    if t is None:
        return None
    return t.data_ptr()

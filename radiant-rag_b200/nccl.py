"""NCCL collectives issued directly on the CALLER'S stream (ctypes binding of libnccl).

The row-sharded step exchanges a few hundred KB of candidate lists; what it pays
is launch + collective latency, so the whole step - kernels AND exchanges - must be one CUDA
graph.  ``torch.distributed``'s NCCL process group runs collectives on its own internal stream
behind event hand-offs and a watchdog thread, which does not capture reliably; ``ncclAllGather``
/ ``ncclAllReduce`` called on the capturing stream do (NCCL >= 2.9 supports stream capture).
This is the ``rr_allgather_candidates(ncclComm_t, ...)`` seam SURVEY.md 8b allows, kept in
Python: the library already loaded by torch is bound with ctypes (no second NCCL in the
process), the unique id travels through the existing ``torch.distributed`` group.
"""

from __future__ import annotations

import ctypes as C
import glob
import os
from typing import Any, Optional

import torch
import torch.distributed as dist

_NCCL_UNIQUE_ID_BYTES = 128


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * _NCCL_UNIQUE_ID_BYTES)]


# ncclDataType_t / ncclRedOp_t (nccl.h)
_DTYPE = {torch.int8: 0, torch.uint8: 1, torch.int32: 2, torch.int64: 4, torch.float16: 6, torch.float32: 7,
          torch.float64: 8, torch.bfloat16: 9}
_SUM, _PROD, _MAX, _MIN = 0, 1, 2, 3


def _find_library() -> str:
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*"))
    cands += glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libnccl.so*"))
    if cands:
        return os.path.realpath(sorted(cands)[0])
    return "libnccl.so.2"


class NcclComm:
    """One communicator over the ranks of `group` (default: the world), created collectively."""

    def __init__(self, device: torch.device, group: Optional[Any] = None) -> None:
        if not dist.is_initialized():
            raise RuntimeError("NcclComm needs torch.distributed to be initialised (one process per GPU)")
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        lib = C.CDLL(_find_library())
        lib.ncclGetErrorString.restype = C.c_char_p
        lib.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
        lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        lib.ncclAllGather.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        lib.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.ncclCommDestroy.argtypes = [C.c_void_p]
        self._lib = lib
        uid = _UniqueId()
        if self.rank == 0:
            self._check(lib.ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
        buf = torch.tensor(list(bytes(uid.internal)), dtype=torch.uint8, device=self.device)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(buf, src=src, group=group)
        raw = bytes(buf.cpu().tolist())
        C.memmove(C.addressof(uid), raw, _NCCL_UNIQUE_ID_BYTES)
        comm = C.c_void_p()
        with torch.cuda.device(self.device):
            self._check(lib.ncclCommInitRank(C.byref(comm), self.world, uid, self.rank), "ncclCommInitRank")
        self._comm = comm
        # first collective connects the channels: do it now, outside any capture
        warm = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.all_reduce(warm, "sum")
        torch.cuda.current_stream(self.device).synchronize()

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what} failed: {self._lib.ncclGetErrorString(rc).decode()}")

    def all_gather(self, send: torch.Tensor, recv: torch.Tensor) -> torch.Tensor:
        """recv [world, *send.shape] <- send of every rank, on the current stream."""
        assert send.is_contiguous() and recv.is_contiguous() and recv.numel() == send.numel() * self.world
        st = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._lib.ncclAllGather(send.data_ptr(), recv.data_ptr(), send.numel(), _DTYPE[send.dtype],
                                            self._comm, st), "ncclAllGather")
        return recv

    def all_reduce(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
        """In-place all-reduce on the current stream (op: sum | max | min)."""
        assert t.is_contiguous()
        code = {"sum": _SUM, "max": _MAX, "min": _MIN}[op]
        st = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._lib.ncclAllReduce(t.data_ptr(), t.data_ptr(), t.numel(), _DTYPE[t.dtype], code,
                                            self._comm, st), "ncclAllReduce")
        return t

    def close(self) -> None:
        if getattr(self, "_comm", None):
            self._lib.ncclCommDestroy(self._comm)
            self._comm = None

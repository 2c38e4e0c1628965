"""Build librr_b200.so in-tree with nvcc for sm_100a (no torch dependency).

Every .cu is compiled to its own object (in parallel, only when stale) and the objects are
linked into one shared library; nvcc cross-compiles without a GPU, so this also runs in the
CPU-only build container."""

from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = PKG_DIR.parent / "build" / "obj"
LIB_PATH = PKG_DIR / "librr_b200.so"
SOURCES = ["api.cu", "quantize.cu", "hamming.cu", "rescore.cu", "exact.cu", "bm25.cu", "bm25_fast.cu",
           "rrf.cu", "synth.cu", "tc_search.cu", "tc_exact.cu", "probe.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _headers():
    return list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "radiant_rag_b200.h"]


def _obj_stale(src: Path, obj: Path) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(d.stat().st_mtime > t for d in [src, Path(__file__)] + _headers())


def build_library(force: bool = False, verbose: bool = False) -> Path:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # kernel-variant experiments: RR_BUILD_TAG=x RR_NVCC_EXTRA="-DRR_BF_WARPS=32" builds
    # librr_b200_x.so from its own object directory (load it with RR_B200_LIB=...)
    extra = os.environ.get("RR_NVCC_EXTRA", "").split()
    tag = os.environ.get("RR_BUILD_TAG", "")
    obj_dir = OBJ_DIR if not tag else OBJ_DIR.parent / f"obj_{tag}"
    lib_path = LIB_PATH if not tag else PKG_DIR / f"librr_b200_{tag}.so"
    obj_dir.mkdir(parents=True, exist_ok=True)
    jobs = []
    for s in SOURCES:
        src, obj = CSRC / s, obj_dir / (s[:-3] + ".o")
        if force or _obj_stale(src, obj):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            jobs.append((s, cmd))

    def run(job):
        name, cmd = job
        res = subprocess.run(cmd, capture_output=True, text=True)
        return name, res

    failed = False
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        for name, res in pool.map(run, jobs):
            if res.returncode != 0:
                sys.stderr.write(f"--- {name}\n{res.stdout}{res.stderr}")
                failed = True
            elif verbose:
                sys.stderr.write(f"--- {name}\n{res.stderr}")
    if failed:
        raise RuntimeError("nvcc failed building librr_b200.so")
    objs = [obj_dir / (s[:-3] + ".o") for s in SOURCES]
    if jobs or not lib_path.exists() or any(o.stat().st_mtime > lib_path.stat().st_mtime for o in objs):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", str(lib_path)] + [str(o) for o in objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed linking librr_b200.so")
    return lib_path


if __name__ == "__main__":
    print(build_library(force="-f" in sys.argv, verbose="-v" in sys.argv))

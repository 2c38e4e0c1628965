"""CPU: librr_b200.so loads without a GPU and exports every symbol the header declares;
the ctypes prototype table lists the same set.  No compute calls here."""

import ctypes
import re
import subprocess
from pathlib import Path

import pytest

import __graft_entry__ as graft
from radiant_rag_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "radiant_rag_b200.h"


@pytest.fixture(scope="module")
def lib():
    graft.build_native()
    return _lib.load()


def header_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_prototype_table_matches_header():
    assert sorted(_lib.PROTOTYPES) == header_symbols()


def test_exports_are_c_linkage(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (rr_[a-z0-9_]+)$", out, flags=re.M))
    assert exported == set(header_symbols())


def test_version_and_error_string(lib):
    assert lib.rr_abi_version() == 2
    assert isinstance(lib.rr_last_error(), bytes)


def test_built_for_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_gpu_means_loud_failure(lib):
    """Without a device the product path raises; it never falls back to the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.RadiantB200Error):
        _lib.init(0)
    from radiant_rag_b200.index import DenseIndex

    with pytest.raises(_lib.RadiantB200Error):
        DenseIndex(64, device=0)
    with pytest.raises(_lib.RadiantB200Error):
        DenseIndex(64, device="cpu")


def test_product_never_imports_oracle():
    pkg = ROOT / "radiant-rag_b200"
    tools = ROOT / "tools"  # measurement tools are not checkers either
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(tools.glob("*.py")):
        text = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), p
        assert "/root/reference" not in text, p

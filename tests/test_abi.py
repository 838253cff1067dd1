"""The C-ABI library loads and exports every symbol include/bellman_b200.h declares (no compute
calls: there is no GPU here), and refuses to run without a CUDA device instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "bellman_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bmpc_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_binding_list():
    from bellman_mpc_b200 import _lib
    assert header_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_symbol():
    from bellman_mpc_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libbellman_b200.so not built (run `make -j8`)")
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_gpu():
    """Without a usable CUDA device context creation must fail loudly (IoError), not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bellman_mpc_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libbellman_b200.so not built")
    import bellman_mpc_b200 as bm
    with pytest.raises(bm.IoError):
        bm.Worker(0)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under bellman_mpc_b200/ may reference it"""
    pkg = os.path.join(ROOT, "bellman_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text.replace("oracle/fields.py", "") or f.endswith((".cuh", ".cu")), f

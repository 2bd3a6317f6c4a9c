"""The C-ABI library loads and exports every symbol include/hpdg_b200.h declares (CPU only; no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "hpdg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hpdg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(hp):
    syms = header_symbols()
    assert len(syms) >= 35
    L = ctypes.CDLL(hp.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/hpdg_b200.h but not exported"
    # the Python mirror binds exactly the declared surface
    assert sorted(hp.SIGNATURES) == syms


def test_no_cpu_fallback(hp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hp.HpdgError, match="no CUDA device"):
        hp.Context((2, 2, 2), degree=1)


def test_create_argument_errors(hp):
    with pytest.raises(hp.HpdgError, match="dim must be 2 or 3"):
        hp.Context((4,), degree=1)
    with pytest.raises(hp.HpdgError, match="degree"):
        hp.Context((2, 2), degree=[1, 2, 3])
    with pytest.raises(hp.HpdgError, match="out of range"):
        hp.Context((2, 2), degree=14)


def test_product_does_not_touch_oracle():
    # the product path must never import, link or call anything under oracle/
    pkg = os.path.join(ROOT, "dune-hpdg_b200")
    for dp, _, files in os.walk(pkg):
        if os.sep + "build" in dp or os.sep + "lib" in dp:
            continue
        for f in files:
            if f.endswith((".cu", ".cc", ".hpp", ".h", ".py", ".hh")) or f == "Makefile":
                txt = open(os.path.join(dp, f)).read()
                assert "hpdg_oracle" not in txt and "orc_" not in txt and "from oracle" not in txt and "import oracle" not in txt, f

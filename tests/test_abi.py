"""The C-ABI library loads and exports every symbol include/hpdg_b200.h declares (CPU only; no compute)."""
import ctypes
import os
import re

import numpy as np
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


def test_face_coupling_tables_reproduce_polynomials(hp):
    # host-only check of the product's tangential face couplings (csrc/tables.cc), no GPU: an L2 projection onto P_pe reproduces
    # polynomials of degree <= pe, so for q of degree <= min(pe, po) the coupling applied to the neighbour's nodal values of q must
    # give this element's nodal values of q -- on conforming faces (variableipdg.hh:326-361) and on the fine / coarse side of a
    # hanging face (sfipdg.hh:472-491: the neighbour sees the intersection through tau_o = (tau + {0,1}) / 2 resp. 2 tau - {0,1}).
    rng = np.random.default_rng(2)
    for pe in range(1, 8):
        xe = hp.tables_1d(pe)[0]
        for po in range(1, 8):
            xo = hp.tables_1d(po)[0]
            q = np.polynomial.Polynomial(rng.standard_normal(min(pe, po) + 1))
            P, _ = hp.tables_face(pe, po, 0)
            assert np.abs(P @ q(xo) - q(xe)).max() < 1e-12
            for half in (0, 1):
                # e fine on the low / high half of o's side: o's coordinate of e's point tau is (tau + half) / 2
                P, _ = hp.tables_face(pe, po, 3 + half)
                assert np.abs(P @ q(xo) - q((xe + half) / 2)).max() < 1e-12
            # e coarse: its two fine neighbours hold q(tau) at tau = (x_o + half) / 2; the two half couplings together project q
            tot = np.zeros(pe + 1)
            own = np.zeros((pe + 1, pe + 1))
            for half in (0, 1):
                P, Q = hp.tables_face(pe, po, 1 + half)
                tot += P @ q((xo + half) / 2)
                own += Q
            assert np.abs(tot - q(xe)).max() < 1e-12
            assert np.abs(own - np.eye(pe + 1)).max() < 1e-12

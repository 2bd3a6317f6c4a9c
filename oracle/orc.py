"""ctypes front end of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path (hpdg_b200) never does.  See hpdg_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libhpdg_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "hpdg_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp = C.c_void_p
    sig = {
        "orc_gl_nodes": (None, [C.c_int, _dp]),
        "orc_gauss_legendre": (None, [C.c_int, _dp, _dp]),
        "orc_lagrange": (C.c_double, [C.c_int, _dp, C.c_int, C.c_double]),
        "orc_lagrange_prime": (C.c_double, [C.c_int, _dp, C.c_int, C.c_double]),
        "orc_mesh_create": (vp, [C.c_int, _ip, _dp, _ip, C.c_void_p, C.c_double, C.c_int]),
        "orc_mesh_destroy": (None, [vp]),
        "orc_mesh_dimension": (C.c_long, [vp]),
        "orc_mesh_nelem": (C.c_long, [vp]),
        "orc_mesh_offsets": (None, [vp, _lp]),
        "orc_interpolate_normsq": (None, [vp, _dp]),
        "orc_apply_mf": (None, [vp, _dp, _dp, C.c_double, C.c_int]),
        "orc_assemble": (vp, [vp]),
        "orc_bcrs_destroy": (None, [vp]),
        "orc_bcrs_nblocks": (C.c_long, [vp]),
        "orc_bcrs_nentries": (C.c_long, [vp]),
        "orc_bcrs_export": (None, [vp, _lp, _ip, _lp, _dp]),
        "orc_bcrs_mv": (None, [vp, _dp, _dp, C.c_int]),
        "orc_bcrs_mmv": (None, [vp, _dp, _dp]),
        "orc_bcrs_frobenius_diff": (C.c_double, [vp, vp]),
        "orc_bcrs_diag_block": (None, [vp, C.c_long, _dp]),
        "orc_blockgs_iterate": (None, [vp, _dp, _dp]),
        "orc_l1_regularization": (None, [vp, _lp, C.c_long, _dp]),
        "orc_l1_iterate": (None, [vp, _dp, _dp, _dp]),
        "orc_blockjacobi_apply": (None, [vp, _dp, _dp, C.c_double, C.c_int]),
        "orc_diag_block_mf": (None, [vp, C.c_long, _dp]),
        "orc_transfer_matrix": (None, [C.c_int, C.c_int, C.c_int, _dp]),
        "orc_mesh_coarsen": (vp, [vp, C.c_int]),
        "orc_restrict": (None, [vp, vp, _dp, _dp]),
        "orc_prolong": (None, [vp, vp, _dp, _dp]),
        "orc_galerkin_restrict": (vp, [vp, vp, vp]),
        "orc_vcycle": (None, [C.c_int, C.POINTER(vp), C.POINTER(vp), C.c_int, C.c_double,
                              C.c_int, C.c_int, C.c_int, _dp, _dp]),
        "orc_fill_random": (None, [_dp, C.c_long, C.c_uint]),
        "orc_max_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def max_threads():
    return lib().orc_max_threads()


def gl_nodes(p):
    x = np.zeros(p + 1)
    lib().orc_gl_nodes(p, x)
    return x


def gauss_legendre(m):
    x = np.zeros(m)
    w = np.zeros(m)
    lib().orc_gauss_legendre(m, x, w)
    return x, w


def lagrange(p, i, x):
    return lib().orc_lagrange(p, gl_nodes(p), i, float(x))


def lagrange_prime(p, i, x):
    return lib().orc_lagrange_prime(p, gl_nodes(p), i, float(x))


def fill_random(n, seed=1887):
    v = np.zeros(n)
    lib().orc_fill_random(v, n, seed)
    return v


def transfer_matrix(dim, pc, pf):
    T = np.zeros(((pf + 1) ** dim, (pc + 1) ** dim))
    lib().orc_transfer_matrix(dim, pc, pf, T)
    return T


class Matrix:
    """DynamicBCRSMatrix-layout block CSR (common/dynamicbcrs.hh:178-199)."""

    def __init__(self, handle, mesh):
        self.h = handle
        self.mesh = mesh

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            try:
                _lib.orc_bcrs_destroy(self.h)
            except Exception:
                pass
            self.h = None

    @property
    def nblocks(self):
        return lib().orc_bcrs_nblocks(self.h)

    @property
    def nentries(self):
        return lib().orc_bcrs_nentries(self.h)

    def export(self):
        nb = self.nblocks
        rowptr = np.zeros(self.mesh.nelem + 1, dtype=np.int64)
        col = np.zeros(nb, dtype=np.int32)
        boff = np.zeros(nb + 1, dtype=np.int64)
        val = np.zeros(self.nentries)
        lib().orc_bcrs_export(self.h, rowptr, col, boff, val)
        return rowptr, col, boff, val

    def mv(self, x, threads=1):
        y = np.zeros_like(x)
        lib().orc_bcrs_mv(self.h, np.ascontiguousarray(x), y, threads)
        return y

    def mmv(self, x, y):
        lib().orc_bcrs_mmv(self.h, np.ascontiguousarray(x), y)
        return y

    def diag_block(self, e):
        n = self.mesh.block_size(e)
        out = np.zeros((n, n))
        lib().orc_bcrs_diag_block(self.h, e, out)
        return out

    def blockgs_iterate(self, b, x):
        lib().orc_blockgs_iterate(self.h, np.ascontiguousarray(b), x)
        return x

    def l1_regularization(self, ghosts):
        """L1Smoother::preprocess (iterationsteps/l1smoother.hh:31-57)."""
        g = np.ascontiguousarray(ghosts, dtype=np.int64)
        reg = np.zeros(self.mesh.ndof)
        lib().orc_l1_regularization(self.h, g, len(g), reg)
        return reg

    def l1_iterate(self, reg, b, x):
        """L1Smoother::iterate (iterationsteps/l1smoother.hh:63-113)."""
        lib().orc_l1_iterate(self.h, np.ascontiguousarray(reg), np.ascontiguousarray(b), x)
        return x

    def frobenius_diff(self, other):
        return lib().orc_bcrs_frobenius_diff(self.h, other.h)

    def to_dense(self):
        rowptr, col, boff, val = self.export()
        off = self.mesh.offsets
        N = off[-1]
        A = np.zeros((N, N))
        for i in range(self.mesh.nelem):
            for k in range(rowptr[i], rowptr[i + 1]):
                j = col[k]
                r, c = off[i + 1] - off[i], off[j + 1] - off[j]
                A[off[i]:off[i + 1], off[j]:off[j + 1]] = val[boff[k]:boff[k] + r * c].reshape(r, c)
        return A


class Mesh:
    """Structured YaspGrid-like mesh with a per-element degree map
    (functionspacebases/dynamicdgqkglbasis.hh:36-197)."""

    def __init__(self, n, L=None, degree=1, sigma=2.0, dirichlet=True, pen_degree=None, _handle=None):
        if _handle is not None:
            self.h = _handle
        self.dim = len(n)
        self.n = np.asarray(n, dtype=np.int32)
        self.L = np.asarray(L if L is not None else [1.0] * self.dim, dtype=np.float64)
        ne = int(np.prod(self.n))
        self.degree = np.ascontiguousarray(np.broadcast_to(np.asarray(degree, dtype=np.int32), (ne,)))
        self.sigma = float(sigma)
        self.dirichlet = bool(dirichlet)
        self.pen_degree = None if pen_degree is None else np.ascontiguousarray(
            np.broadcast_to(np.asarray(pen_degree, dtype=np.int32), (ne,)))
        if _handle is None:
            pd = None if self.pen_degree is None else self.pen_degree.ctypes.data_as(C.c_void_p)
            self.h = lib().orc_mesh_create(self.dim, self.n, self.L, self.degree, pd, self.sigma,
                                           int(self.dirichlet))
        self.nelem = lib().orc_mesh_nelem(self.h)
        self.ndof = lib().orc_mesh_dimension(self.h)
        self.offsets = np.zeros(self.nelem + 1, dtype=np.int64)
        lib().orc_mesh_offsets(self.h, self.offsets)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            try:
                _lib.orc_mesh_destroy(self.h)
            except Exception:
                pass
            self.h = None

    def block_size(self, e):
        return int(self.offsets[e + 1] - self.offsets[e])

    def interpolate_normsq(self):
        x = np.zeros(self.ndof)
        lib().orc_interpolate_normsq(self.h, x)
        return x

    def apply_mf(self, x, factor=1.0, threads=1):
        y = np.zeros(self.ndof)
        lib().orc_apply_mf(self.h, np.ascontiguousarray(x), y, factor, threads)
        return y

    def assemble(self):
        return Matrix(lib().orc_assemble(self.h), self)

    def blockjacobi_apply(self, r, factor=1.0, local_solver=0):
        c = np.zeros(self.ndof)
        lib().orc_blockjacobi_apply(self.h, np.ascontiguousarray(r), c, factor, local_solver)
        return c

    def diag_block_mf(self, e):
        n = self.block_size(e)
        out = np.zeros((n, n))
        lib().orc_diag_block_mf(self.h, e, out)
        return out

    def coarsen(self, max_order):
        pen = self.pen_degree if self.pen_degree is not None else self.degree
        return Mesh(self.n, self.L, np.minimum(self.degree, max_order), self.sigma, self.dirichlet,
                    pen_degree=pen)

    def restrict(self, coarse, xf):
        xc = np.zeros(coarse.ndof)
        lib().orc_restrict(self.h, coarse.h, np.ascontiguousarray(xf), xc)
        return xc

    def prolong(self, coarse, xc):
        xf = np.zeros(self.ndof)
        lib().orc_prolong(self.h, coarse.h, np.ascontiguousarray(xc), xf)
        return xf

    def galerkin_restrict(self, coarse, Af):
        return Matrix(lib().orc_galerkin_restrict(self.h, coarse.h, Af.h), coarse)


def vcycle(levels, mats, x, b, smoother=0, damping=1.0, pre=5, post=5, coarse_its=5):
    """One Multigrid<Vector>::apply (iterationsteps/mg/multigrid_impl.hh:16-61). levels[0] is the
    coarsest.  Returns (x_new, residual)."""
    n = len(levels)
    LT = C.c_void_p * n
    lv = LT(*[m.h for m in levels])
    mt = LT(*[m.h for m in mats]) if mats is not None else None
    x = np.array(x, dtype=np.float64)
    b = np.array(b, dtype=np.float64)
    lib().orc_vcycle(n, lv, mt, smoother, damping, pre, post, coarse_its, x, b)
    return x, b

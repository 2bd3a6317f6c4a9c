"""Python host-side mirror of the reference interface over the C ABI (include/hpdg_b200.h).

This is the ctypes stub a maintainer would add on the reference side (INTEGRATION.md shows the C++
one); it carries no arithmetic.  Names follow the reference: `Operator.apply(x, Ax)`
(matrix-free/operator.hh:41), `BlockJacobi` as a `Smoother(c, r)` (iterationsteps/mg/multigrid.hh:13-14),
`OrderTransfer.restrict/prolong` (transferoperators/ordertransfer.hh:91-119), `Multigrid.apply(x, b)`
(iterationsteps/mg/multigrid_impl.hh:16).  The library has NO CPU fallback: constructing a Context
without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HPDG_B200_LIB: load another build of the same library (kernel A/B experiments of tools/, e.g. `make EXTRA=-DQ3P_...`)
LIB_PATH = os.environ.get("HPDG_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libhpdg_b200.so")

FINEST = -1
JACOBI_DENSE = 0
JACOBI_FD = 1
SMOOTHER_BLOCKGS = 2
SMOOTHER_BLOCKGS_MF = 3
PRECOND_NONE = 0
PRECOND_JACOBI = 1
PRECOND_VCYCLE = 2

_lib = None
_vp = C.c_void_p
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")

# every symbol include/hpdg_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hpdg_create": (C.c_int, [C.POINTER(_vp), C.c_int, _ip, _dp, _ip, C.c_long, C.c_double, C.c_int, C.c_int]),
    "hpdg_create_distributed": (C.c_int, [C.POINTER(_vp), C.c_int, _ip, _dp, C.c_int, C.c_double, C.c_int, C.c_int,
                                          _ip, C.c_int, C.c_int, C.c_void_p]),
    "hpdg_create_distributed_hp": (C.c_int, [C.POINTER(_vp), C.c_int, _ip, _dp, _ip, C.c_double, C.c_int, C.c_int,
                                             _ip, C.c_int, C.c_int, C.c_void_p]),
    "hpdg_create_refined_2d": (C.c_int, [C.POINTER(_vp), _ip, _dp, C.c_void_p, _ip, C.c_long, C.c_double, C.c_int, C.c_int]),
    "hpdg_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "hpdg_halo_ipc_handle": (C.c_int, [_vp, C.c_void_p]),
    "hpdg_halo_ipc_attach": (C.c_int, [_vp, C.c_void_p]),
    "hpdg_destroy": (None, [_vp]),
    "hpdg_last_error": (C.c_char_p, [_vp]),
    "hpdg_set_option": (C.c_int, [_vp, C.c_char_p, C.c_long]),
    "hpdg_num_levels": (C.c_int, [_vp]),
    "hpdg_num_elements": (C.c_long, [_vp]),
    "hpdg_dimension": (C.c_long, [_vp, C.c_int]),
    "hpdg_block_offsets": (C.c_int, [_vp, C.c_int, _lp]),
    "hpdg_level_degrees": (C.c_int, [_vp, C.c_int, _ip]),
    "hpdg_build_p_hierarchy": (C.c_int, [_vp]),
    "hpdg_vec_alloc": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "hpdg_vec_free": (C.c_int, [_vp, _vp]),
    "hpdg_vec_upload": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_vec_download": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_host_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "hpdg_host_free": (C.c_int, [_vp, _vp]),
    "hpdg_sync": (C.c_int, [_vp]),
    "hpdg_stream": (_vp, [_vp]),
    "hpdg_op_apply": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_double]),
    "hpdg_op_apply_device": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_double]),
    "hpdg_op_apply_async": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_double]),
    "hpdg_op_apply_accum": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_double]),
    "hpdg_op_apply_accum_device": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_double]),
    "hpdg_op_apply_accum_async": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_double]),
    "hpdg_jacobi_setup": (C.c_int, [_vp, C.c_int, C.c_int]),
    "hpdg_jacobi_apply": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_double]),
    "hpdg_jacobi_apply_device": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_double]),
    "hpdg_jacobi_apply_async": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_double]),
    "hpdg_jacobi_bytes": (C.c_size_t, [_vp, C.c_int, C.c_int]),
    "hpdg_diag_block": (C.c_int, [_vp, C.c_int, C.c_long, _dp]),
    "hpdg_bcrs_sizes": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_long)]),
    "hpdg_assemble_bcrs": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "hpdg_bcrs_mv": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_bcrs_mv_device": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_blockgs_iterate": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_blockgs_iterate_device": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_blockgs_mf_iterate": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_blockgs_mf_iterate_device": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_l1_setup": (C.c_int, [_vp, C.c_int, _vp, C.c_long]),
    "hpdg_l1_iterate": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_l1_iterate_device": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_restrict": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_prolong": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_restrict_device": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_prolong_device": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_vcycle": (C.c_int, [_vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "hpdg_vcycle_device": (C.c_int, [_vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "hpdg_dot_device": (C.c_int, [_vp, C.c_int, _vp, _vp, C.POINTER(C.c_double)]),
    "hpdg_two_norm_device": (C.c_int, [_vp, C.c_int, _vp, C.POINTER(C.c_double)]),
    "hpdg_axpy_device": (C.c_int, [_vp, C.c_int, C.c_double, _vp, _vp]),
    "hpdg_scale_device": (C.c_int, [_vp, C.c_int, C.c_double, _vp]),
    "hpdg_assign_device": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "hpdg_pcg": (C.c_int, [_vp, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, C.c_double, C.c_int, C.c_int,
                           C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "hpdg_pcg_device": (C.c_int, [_vp, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, C.c_double, C.c_int, C.c_int,
                                  C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "hpdg_loop_solve_device": (C.c_int, [_vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_double, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "hpdg_tables_1d": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "hpdg_tables_face": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "hpdg_launch_count": (C.c_long, [_vp]),
    "hpdg_uses_uniform_kernel": (C.c_int, [_vp, C.c_int]),
    "hpdg_time_apply_device": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.POINTER(C.c_float)]),
    "hpdg_time_jacobi_device": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_double, C.c_int, C.POINTER(C.c_float)]),
}


def tables_1d(p):
    """(nodes, M, S, t, g) of degree p as the kernels use them (host-only introspection)"""
    n = p + 1
    nodes, M, S, t, g = np.zeros(n), np.zeros((n, n)), np.zeros((n, n)), np.zeros((2, n)), np.zeros((2, n))
    if lib().hpdg_tables_1d(p, nodes.ctypes.data, M.ctypes.data, S.ctypes.data, t.ctypes.data, g.ctypes.data):
        raise HpdgError(lib().hpdg_last_error(None).decode())
    return nodes, M, S, t, g


def tables_face(pe, po, kind):
    """(coupling (pe+1) x (po+1), own-side coupling or None) of a conforming (kind 0) / hanging (kinds 1..4) face (host only)"""
    out = np.zeros((pe + 1, po + 1))
    own = np.zeros((pe + 1, pe + 1)) if kind in (1, 2) else None
    if lib().hpdg_tables_face(pe, po, kind, out.ctypes.data, own.ctypes.data if own is not None else None):
        raise HpdgError(lib().hpdg_last_error(None).decode())
    return out, own


class HpdgError(RuntimeError):
    """Raised where the reference would DUNE_THROW (e.g. iterationsteps/dynamicblockgs.hh:117)."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HpdgError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _hptr(a):
    """host pointer of a numpy array or a raw integer address"""
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return int(a)


class Context:
    """One GPU's view of the problem: mesh brick + degree map + penalty (replaces the reference's
    basis + IPDGOperator constructor arguments)."""

    def __init__(self, n, L=None, degree=1, sigma=2.0, dirichlet=True, device=0, pgrid=None, rank=0, nranks=1,
                 nccl_id=None):
        self._h = _vp()
        self.dim = len(n)
        n = np.ascontiguousarray(n, dtype=np.int32)
        L = np.ascontiguousarray(L if L is not None else [1.0] * self.dim, dtype=np.float64)
        deg = np.ascontiguousarray(np.atleast_1d(degree), dtype=np.int32)
        if pgrid is None:
            rc = lib().hpdg_create(C.byref(self._h), self.dim, n, L, deg, deg.size, sigma, int(dirichlet), device)
        else:
            pg = np.ascontiguousarray(pgrid, dtype=np.int32)
            idp = C.c_char_p(nccl_id) if nccl_id is not None else None
            if deg.size == 1:
                rc = lib().hpdg_create_distributed(C.byref(self._h), self.dim, n, L, int(deg[0]), sigma, int(dirichlet),
                                                   device, pg, rank, nranks, idp)
            else:  # per-element degree map of the local brick: distributed hp
                rc = lib().hpdg_create_distributed_hp(C.byref(self._h), self.dim, n, L, deg, sigma, int(dirichlet),
                                                      device, pg, rank, nranks, idp)
        if rc:
            msg = lib().hpdg_last_error(None).decode()
            self._h = None
            raise HpdgError(msg)
        self.n = n

    @classmethod
    def refined_2d(cls, n, refine, degree, L=None, sigma=2.0, dirichlet=True, device=0):
        """Non-conforming 2-D mesh: the base grid `n` with the cells flagged in `refine` (x fastest) split once into 2 x 2 children
        (hpdg_create_refined_2d); `degree` holds one entry per leaf element (base-cell order, children x fastest)."""
        self = cls.__new__(cls)
        self._h = _vp()
        self.dim = 2
        nn = np.ascontiguousarray(n, dtype=np.int32)
        LL = np.ascontiguousarray(L if L is not None else [1.0, 1.0], dtype=np.float64)
        rf = np.ascontiguousarray(refine, dtype=np.uint8)
        assert rf.size == int(nn[0]) * int(nn[1])
        nleaf = int(rf.size + 3 * np.count_nonzero(rf))
        deg = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(degree), (nleaf,)) if np.size(degree) == 1 else degree, dtype=np.int32)
        rc = lib().hpdg_create_refined_2d(C.byref(self._h), nn, LL, rf.ctypes.data_as(C.c_void_p), deg, deg.size, sigma,
                                          int(dirichlet), device)
        if rc:
            msg = lib().hpdg_last_error(None).decode()
            self._h = None
            raise HpdgError(msg)
        self.n = nn
        return self

    def halo_ipc_handle(self):
        buf = C.create_string_buffer(64)
        self._ck(lib().hpdg_halo_ipc_handle(self._h, buf))
        return buf.raw

    def halo_ipc_attach(self, handles_by_rank):
        blob = b"".join(handles_by_rank)
        self._ck(lib().hpdg_halo_ipc_attach(self._h, C.c_char_p(blob)))

    def _ck(self, rc):
        if rc:
            raise HpdgError(lib().hpdg_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().hpdg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # sizes ------------------------------------------------------------------------------------
    @property
    def num_levels(self):
        return lib().hpdg_num_levels(self._h)

    @property
    def num_elements(self):
        return lib().hpdg_num_elements(self._h)

    def dimension(self, level=FINEST):
        n = lib().hpdg_dimension(self._h, level)
        if n < 0:
            raise HpdgError(lib().hpdg_last_error(self._h).decode())
        return n

    def block_offsets(self, level=FINEST):
        off = np.zeros(self.num_elements + 1, dtype=np.int64)
        self._ck(lib().hpdg_block_offsets(self._h, level, off))
        return off

    def level_degrees(self, level=FINEST):
        d = np.zeros(self.num_elements, dtype=np.int32)
        self._ck(lib().hpdg_level_degrees(self._h, level, d))
        return d

    def build_p_hierarchy(self):
        self._ck(lib().hpdg_build_p_hierarchy(self._h))
        return self.num_levels

    def set_option(self, name, value):
        self._ck(lib().hpdg_set_option(self._h, name.encode(), int(value)))

    # device vectors ---------------------------------------------------------------------------
    def vec_alloc(self, level=FINEST):
        p = _vp()
        self._ck(lib().hpdg_vec_alloc(self._h, level, C.byref(p)))
        return p.value

    def vec_free(self, d):
        self._ck(lib().hpdg_vec_free(self._h, d))

    def upload(self, h, d=None, level=FINEST):
        if d is None:
            d = self.vec_alloc(level)
        self._ck(lib().hpdg_vec_upload(self._h, level, _hptr(h), d))
        return d

    def download(self, d, level=FINEST, out=None):
        if out is None:
            out = np.zeros(self.dimension(level))
        self._ck(lib().hpdg_vec_download(self._h, level, d, _hptr(out)))
        return out

    def host_alloc(self, ndoubles):
        """pinned host array (numpy view)"""
        p = _vp()
        self._ck(lib().hpdg_host_alloc(self._h, ndoubles * 8, C.byref(p)))
        buf = (C.c_double * ndoubles).from_address(p.value)
        a = np.frombuffer(buf, dtype=np.float64)
        self.__dict__.setdefault('_pinned', []).append(buf)
        return a, p.value

    def host_free(self, p):
        self._ck(lib().hpdg_host_free(self._h, p))

    def sync(self):
        self._ck(lib().hpdg_sync(self._h))

    @property
    def stream(self):
        return lib().hpdg_stream(self._h)

    @property
    def launch_count(self):
        return lib().hpdg_launch_count(self._h)

    def uses_uniform_kernel(self, level=FINEST):
        return bool(lib().hpdg_uses_uniform_kernel(self._h, level))

    def dot_device(self, dx, dy, level=FINEST):
        r = C.c_double()
        self._ck(lib().hpdg_dot_device(self._h, level, dx, dy, C.byref(r)))
        return r.value

    def axpy_device(self, a, dx, dy, level=FINEST):
        self._ck(lib().hpdg_axpy_device(self._h, level, a, dx, dy))

    def two_norm_device(self, dx, level=FINEST):
        r = C.c_double()
        self._ck(lib().hpdg_two_norm_device(self._h, level, dx, C.byref(r)))
        return r.value

    def scale_device(self, a, dx, level=FINEST):
        self._ck(lib().hpdg_scale_device(self._h, level, a, dx))

    def assign_device(self, dsrc, ddst, level=FINEST):
        self._ck(lib().hpdg_assign_device(self._h, level, dsrc, ddst))

    def time_apply_device(self, dx, dy, reps, level=FINEST):
        ms = C.c_float()
        self._ck(lib().hpdg_time_apply_device(self._h, level, dx, dy, reps, C.byref(ms)))
        return ms.value


class IPDGOperator:
    """The local operator of the tuple `Operator` iterates over: carries its own factor
    (matrix-free/localoperators/localoperator.hh:41-49: factor() / setFactor())."""

    def __init__(self, ctx, level=FINEST, factor=1.0):
        self.ctx, self.level, self._factor = ctx, level, factor

    def factor(self):
        return self._factor

    def setFactor(self, f):
        self._factor = f


class Operator:
    """`Operator::apply(x, Ax)` over a tuple of local operators (matrix-free/operator.hh:41-56): Ax is zeroed once and
    every local operator adds factor_k * (A x).  `Operator(ctx, level, factor)` is the single-operator shorthand;
    `Operator.from_local_operators([op1, op2, ...])` is the tuple form (matrix-free/test/testoperator.cc:80-98)."""

    def __init__(self, ctx, level=FINEST, factor=1.0):
        self.ctx, self.level = ctx, level
        self.local_operators = [IPDGOperator(ctx, level, factor)]

    @classmethod
    def from_local_operators(cls, ops):
        self = cls(ops[0].ctx, ops[0].level, ops[0].factor())
        self.local_operators = list(ops)
        assert all(o.ctx is self.ctx and o.level == self.level for o in ops), "local operators must share context and level"
        return self

    def factor(self):
        return self.local_operators[0].factor()

    def setFactor(self, f):
        self.local_operators[0].setFactor(f)

    def apply(self, x, Ax=None):
        if Ax is None:
            Ax = np.zeros(self.ctx.dimension(self.level))
        for k, op in enumerate(self.local_operators):
            f = lib().hpdg_op_apply if k == 0 else lib().hpdg_op_apply_accum
            self.ctx._ck(f(self.ctx._h, self.level, _hptr(x), _hptr(Ax), op.factor()))
        return Ax

    def apply_device(self, dx, dy, sync=True):
        for k, op in enumerate(self.local_operators):
            last = k == len(self.local_operators) - 1
            if k == 0:
                f = lib().hpdg_op_apply_device if (sync and last) else lib().hpdg_op_apply_async
            else:
                f = lib().hpdg_op_apply_accum_device if (sync and last) else lib().hpdg_op_apply_accum_async
            self.ctx._ck(f(self.ctx._h, self.level, dx, dy, op.factor()))


class BlockJacobi:
    """Smoother<V>(c, r): c = damping * sum_e P_e^T D_e^-1 P_e r (ipdgblockjacobi.hh:58-178)."""

    def __init__(self, ctx, level=FINEST, form=JACOBI_DENSE, damping=1.0):
        self.ctx, self.level, self.form, self.damping = ctx, level, form, damping
        ctx._ck(lib().hpdg_jacobi_setup(ctx._h, level, form))

    def __call__(self, r, c=None):
        if c is None:
            c = np.zeros(self.ctx.dimension(self.level))
        self.ctx._ck(lib().hpdg_jacobi_apply(self.ctx._h, self.level, self.form, _hptr(r), _hptr(c), self.damping))
        return c

    def apply_device(self, dr, dc, sync=True):
        f = lib().hpdg_jacobi_apply_device if sync else lib().hpdg_jacobi_apply_async
        self.ctx._ck(f(self.ctx._h, self.level, self.form, dr, dc, self.damping))

    def time_device(self, dr, dc, reps):
        """ms per application, CUDA events around `reps` back-to-back launches on the context stream"""
        ms = C.c_float()
        self.ctx._ck(lib().hpdg_time_jacobi_device(self.ctx._h, self.level, self.form, dr, dc, self.damping, reps, C.byref(ms)))
        return ms.value

    @property
    def bytes(self):
        return lib().hpdg_jacobi_bytes(self.ctx._h, self.level, self.form)

    def diag_block(self, e):
        off = self.ctx.block_offsets(self.level)
        n = int(off[e + 1] - off[e])
        out = np.zeros((n, n))
        self.ctx._ck(lib().hpdg_diag_block(self.ctx._h, self.level, e, out))
        return out


class AssembledMatrix:
    """The SIPG matrix in DynamicBCRSMatrix layout (common/dynamicbcrs.hh:178-199), assembled on the device:
    the drop-in for `BuildingBlocks::laplace` (buildingblocks/matrices.hh:29-89) + `mv` (matrixwindow.hh:196-209)."""

    def __init__(self, ctx, level=FINEST):
        self.ctx, self.level = ctx, level
        nb, ne = C.c_long(), C.c_long()
        ctx._ck(lib().hpdg_bcrs_sizes(ctx._h, level, C.byref(nb), C.byref(ne)))
        self.block_row_ptr = np.zeros(ctx.num_elements + 1, dtype=np.int64)
        self.block_col = np.zeros(nb.value, dtype=np.int32)
        self.block_off = np.zeros(nb.value + 1, dtype=np.int64)
        self.entries = np.zeros(ne.value)
        ctx._ck(lib().hpdg_assemble_bcrs(ctx._h, level, self.block_row_ptr.ctypes.data, self.block_col.ctypes.data,
                                         self.block_off.ctypes.data, self.entries.ctypes.data))

    def mv(self, x, y=None):
        if y is None:
            y = np.zeros(self.ctx.dimension(self.level))
        self.ctx._ck(lib().hpdg_bcrs_mv(self.ctx._h, self.level, _hptr(x), _hptr(y)))
        return y


class DynamicBlockGS:
    """`LinearIterationStep` face of the reference's default smoother (iterationsteps/dynamicblockgs.hh:87-127)."""

    def __init__(self, matrix):
        self.mat_ = matrix
        self.x_ = self.rhs_ = None

    def setProblem(self, x, rhs):
        self.x_, self.rhs_ = x, rhs

    def preprocess(self):
        pass

    def iterate(self):
        c = self.mat_.ctx
        c._ck(lib().hpdg_blockgs_iterate(c._h, self.mat_.level, _hptr(self.rhs_), _hptr(self.x_)))


class MatrixFreeBlockGS:
    """The same iteration step without an assembled matrix (hpdg_blockgs_mf_iterate): setProblem(x, rhs); iterate().
    `ctx` takes the place of the matrix; works at any mesh size the vectors fit."""

    def __init__(self, ctx, level=FINEST):
        self.ctx, self.level = ctx, level
        self.x_ = self.rhs_ = None

    def setProblem(self, x, rhs):
        self.x_, self.rhs_ = x, rhs

    def preprocess(self):
        pass

    def iterate(self):
        self.ctx._ck(lib().hpdg_blockgs_mf_iterate(self.ctx._h, self.level, _hptr(self.rhs_), _hptr(self.x_)))

    def iterate_device(self, d_x, d_rhs):
        self.ctx._ck(lib().hpdg_blockgs_mf_iterate_device(self.ctx._h, self.level, d_rhs, d_x))


class L1Smoother:
    """`LinearIterationStep` face of the reference's smoother for MPI runs (iterationsteps/l1smoother.hh:20-145):
    L1Smoother(ghosts); setProblem(mat, x, rhs); preprocess(); iterate()."""

    def __init__(self, ghosts):
        self.ghosts_ = np.ascontiguousarray(ghosts, dtype=np.int64)
        self.mat_ = self.x_ = self.rhs_ = None

    def setProblem(self, matrix, x, rhs):
        self.mat_, self.x_, self.rhs_ = matrix, x, rhs

    def preprocess(self):
        c = self.mat_.ctx
        c._ck(lib().hpdg_l1_setup(c._h, self.mat_.level, self.ghosts_.ctypes.data_as(C.c_void_p), len(self.ghosts_)))

    def iterate(self):
        c = self.mat_.ctx
        c._ck(lib().hpdg_l1_iterate(c._h, self.mat_.level, _hptr(self.rhs_), _hptr(self.x_)))


class OrderTransfer:
    """DGOrderTransfer between `fine_level` and `fine_level - 1` (ordertransfer.hh:91-119)."""

    def __init__(self, ctx, fine_level):
        self.ctx = ctx
        self.fine = fine_level if fine_level != FINEST else ctx.num_levels - 1

    def restrict(self, fine_vec, coarse_vec=None):
        if coarse_vec is None:
            coarse_vec = np.zeros(self.ctx.dimension(self.fine - 1))
        self.ctx._ck(lib().hpdg_restrict(self.ctx._h, self.fine, _hptr(fine_vec), _hptr(coarse_vec)))
        return coarse_vec

    def prolong(self, coarse_vec, fine_vec=None):
        if fine_vec is None:
            fine_vec = np.zeros(self.ctx.dimension(self.fine))
        self.ctx._ck(lib().hpdg_prolong(self.ctx._h, self.fine, _hptr(coarse_vec), _hptr(fine_vec)))
        return fine_vec


class Multigrid:
    """Multigrid<Vector>::apply(x, b) (mg/multigrid_impl.hh:16-61) with block-Jacobi smoothing."""

    def __init__(self, ctx, form=JACOBI_FD, damping=0.75, pre=5, post=5, coarse_its=5):
        self.ctx, self.form, self.damping, self.pre, self.post, self.coarse_its = ctx, form, damping, pre, post, coarse_its

    def apply(self, x, b):
        self.ctx._ck(lib().hpdg_vcycle(self.ctx._h, self.form, self.damping, self.pre, self.post, self.coarse_its,
                                       _hptr(x), _hptr(b)))
        return x, b

    def apply_device(self, dx, db):
        self.ctx._ck(lib().hpdg_vcycle_device(self.ctx._h, self.form, self.damping, self.pre, self.post,
                                              self.coarse_its, dx, db))


class LoopSolver:
    """Dune::Solvers::LoopSolver around the multigrid step with the energy norm (buildingblocks/solve.hh:150-166)."""

    def __init__(self, mg, maxIterations=100, tolerance=1e-8):
        self.mg, self.maxIterations, self.tolerance = mg, maxIterations, tolerance
        self.iterationCount_, self.error_ = 0, float("nan")

    def solve_device(self, dx, db):
        it, err = C.c_int(), C.c_double()
        m = self.mg
        m.ctx._ck(lib().hpdg_loop_solve_device(m.ctx._h, m.form, m.damping, m.pre, m.post, m.coarse_its, dx, db,
                                               self.tolerance, self.maxIterations, C.byref(it), C.byref(err)))
        self.iterationCount_, self.error_ = it.value, err.value
        return it.value

    def iterationCount(self):
        return self.iterationCount_


class ConjugateGradients:
    """Preconditioned CG resident on the device (the Krylov loop whose dot products end in the NCCL all-reduce)."""

    def __init__(self, ctx, precond=PRECOND_JACOBI, damping=1.0, smooth=2, coarse_its=5, tol=1e-8, maxit=200, check_every=1):
        self.ctx, self.precond, self.damping, self.smooth, self.coarse_its = ctx, precond, damping, smooth, coarse_its
        self.tol, self.maxit, self.check_every = tol, maxit, check_every
        self.iterations, self.relres = 0, float("nan")

    def _call(self, f, x, b):
        it, rr = C.c_int(), C.c_double()
        self.ctx._ck(f(self.ctx._h, self.precond, self.damping, self.smooth, self.coarse_its, x, b, self.tol, self.maxit,
                       self.check_every, C.byref(it), C.byref(rr)))
        self.iterations, self.relres = it.value, rr.value
        return it.value

    def solve(self, x, b):
        return self._call(lib().hpdg_pcg, _hptr(x), _hptr(b))

    def solve_device(self, dx, db):
        return self._call(lib().hpdg_pcg_device, dx, db)

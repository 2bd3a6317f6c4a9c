"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Bar: relative L2 <= 1e-12 (FP64)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("dirichlet", [True, False])
def test_uniform_3d_vs_assembled(orc, hp, p, dirichlet):
    # ragged extents (not multiples of the 4x4x4 tile), anisotropic spacing
    n, L = (5, 6, 3), [1.0, 1.5, 0.5]
    m = orc.Mesh(n, L=L, degree=p, sigma=2.0, dirichlet=dirichlet)
    x = orc.fill_random(m.ndof)
    ref = m.assemble().mv(x)
    ctx = hp.Context(n, L=L, degree=p, sigma=2.0, dirichlet=dirichlet)
    assert ctx.dimension() == m.ndof and ctx.uses_uniform_kernel()
    y = hp.Operator(ctx).apply(x)
    assert rel(y, ref) < TOL
    ctx.set_option("force_generic", 1)
    assert not ctx.uses_uniform_kernel()
    yg = hp.Operator(ctx, factor=0.5).apply(x)
    assert rel(yg, 0.5 * ref) < TOL


@pytest.mark.parametrize("n", [(1, 1, 1), (2, 1, 3), (4, 4, 4), (8, 8, 8), (9, 4, 5)])
def test_uniform_q3_edge_meshes(orc, hp, n):
    m = orc.Mesh(n, degree=3, sigma=2.0, dirichlet=True)
    x = orc.fill_random(m.ndof)
    ref = m.apply_mf(x, threads=orc.max_threads())
    y = hp.Operator(hp.Context(n, degree=3)).apply(x)
    assert rel(y, ref) < TOL


@pytest.mark.parametrize("n,L", [((8, 4, 12), [1.0, 0.5, 2.0]), ((16, 8, 8), [1.0, 1.0, 1.0]), ((4, 4, 4), [1.0, 1.0, 1.0])])
@pytest.mark.parametrize("dirichlet", [True, False])
@pytest.mark.parametrize("grid", [0, 3])
def test_q3_persistent_kernel(orc, hp, n, L, dirichlet, grid):
    # extents that are multiples of 4 take the persistent bulk-prefetch kernel; grid=3 forces several tiles per CTA
    # (the prefetch loop)
    m = orc.Mesh(n, L=L, degree=3, sigma=2.0, dirichlet=dirichlet)
    x = orc.fill_random(m.ndof)
    ref = m.apply_mf(x, threads=orc.max_threads())
    ctx = hp.Context(n, L=L, degree=3, sigma=2.0, dirichlet=dirichlet)
    ctx.set_option("q3p_grid", grid)
    op = hp.Operator(ctx, factor=-0.75)
    y = op.apply(x)
    assert rel(y, -0.75 * ref) < TOL
    ctx.set_option("variant", 40)  # the one-tile-per-CTA kernel
    assert rel(op.apply(x), y) < TOL


def test_q3_persistent_many_tiles(orc, hp):
    # 32^3: 512 tiles > one per SM slot, so CTAs loop; host-pointer entry point (z-slab launches with a tile offset)
    n = (32, 32, 32)
    m = orc.Mesh(n, degree=3)
    x = orc.fill_random(m.ndof)
    ref = m.apply_mf(x, threads=orc.max_threads())
    ctx = hp.Context(n, degree=3)
    assert rel(hp.Operator(ctx).apply(x), ref) < TOL
    hx, px = ctx.host_alloc(m.ndof)
    hy, py = ctx.host_alloc(m.ndof)
    hx[:] = x
    hp.Operator(ctx, factor=2.0).apply(px, py)
    assert rel(hy, 2.0 * ref) < TOL


def test_cfg1_2d_q2_testdg(orc, hp):
    # BASELINE config 1 / matrix-free/test/testdg.cc: 16x16 Q2, x = |x|^2, factor 0.5, energy error < 1e-14
    m = orc.Mesh((16, 16), degree=2, sigma=2.0, dirichlet=True)
    A = m.assemble()
    x = m.interpolate_normsq()
    ctx = hp.Context((16, 16), degree=2, sigma=2.0, dirichlet=True)
    Ax = hp.Operator(ctx, factor=0.5).apply(x) / 0.5
    d = Ax - A.mv(x)
    assert 0 <= d @ A.mv(d) < 1e-14
    assert rel(Ax, A.mv(x)) < TOL


@pytest.mark.parametrize("dirichlet", [True, False])
def test_2d_against_gauss_lobatto_sumfactorised_formulation(orc, hp, dirichlet):
    # the CUDA 2-D path against the oracle's THIRD formulation (oracle/sf2d.py: restatement of SumFactIPDGOperator,
    # matrix-free/localoperators/sfipdg.hh, Gauss-Lobatto quadrature, own node generator): cfg1's mesh and an hp mesh
    from oracle import sf2d
    x = orc.Mesh((16, 16), degree=2).interpolate_normsq()
    y = hp.Operator(hp.Context((16, 16), degree=2, sigma=2.0, dirichlet=dirichlet), factor=0.5).apply(x)
    assert rel(y, sf2d.SumFactIPDG2D((16, 16), (1.0, 1.0), 2, 2.0, dirichlet).apply(x, 0.5)) < TOL
    rng = np.random.default_rng(5)
    deg = rng.integers(1, 7, 42).astype(np.int32)
    s = sf2d.SumFactIPDG2D((7, 6), (1.0, 2.0), deg, 2.0, dirichlet)
    x = orc.fill_random(s.ndof)
    y = hp.Operator(hp.Context((7, 6), L=[1.0, 2.0], degree=deg, sigma=2.0, dirichlet=dirichlet)).apply(x)
    assert rel(y, s.apply(x)) < TOL


@pytest.mark.parametrize("dirichlet", [True, False])
def test_nonconforming_hanging_node_mesh(orc, hp, dirichlet):
    # f-4: non-conforming faces (matrix-free/localoperators/sfipdg.hh:213-222,472-491).  A base grid with some cells split once
    # (hanging nodes between refined and unrefined cells), per-leaf degree 1..5, against the oracle's restatement of the
    # reference's non-conforming branch (oracle/sf2d.py: RefinedSumFactIPDG2D); also uniform degree, a single refined cell
    # (testdgrestrict.cc-style one-refinement mesh), everything refined, and the entry points that must refuse such a mesh.
    from oracle import sf2d
    rng = np.random.default_rng(17)
    cases = []
    ref = np.zeros(20, dtype=np.uint8); ref[[1, 6, 7, 12, 18]] = 1
    cases.append(((5, 4), ref, (1.0, 1.5), None))
    one = np.zeros(9, dtype=np.uint8); one[4] = 1
    cases.append(((3, 3), one, (1.0, 1.0), 2))
    cases.append(((2, 3), np.ones(6, dtype=np.uint8), (2.0, 1.0), None))
    edge = np.zeros(8, dtype=np.uint8); edge[[0, 3, 4]] = 1      # refined cells on the domain boundary and in a corner
    cases.append(((4, 2), edge, (1.0, 1.0), 3))
    for n, ref, L, p in cases:
        nleaf = int(ref.size + 3 * ref.sum())
        deg = rng.integers(1, 6, nleaf).astype(np.int32) if p is None else np.full(nleaf, p, dtype=np.int32)
        o = sf2d.RefinedSumFactIPDG2D(n, ref, deg, L, 2.0, dirichlet)
        ctx = hp.Context.refined_2d(n, ref, deg, L=L, sigma=2.0, dirichlet=dirichlet)
        assert ctx.dimension() == o.ndof and np.array_equal(ctx.block_offsets(), o.off)
        x = rng.standard_normal(o.ndof)
        y = hp.Operator(ctx, factor=0.5).apply(x)
        assert rel(y, o.apply(x, 0.5)) < TOL
        with pytest.raises(hp.HpdgError):
            hp.BlockJacobi(ctx, form=hp.JACOBI_DENSE)
        with pytest.raises(hp.HpdgError):
            ctx.build_p_hierarchy()
        ctx.close()


@pytest.mark.parametrize("dim,n", [(2, (7, 5)), (3, (4, 3, 5))])
@pytest.mark.parametrize("dirichlet", [True, False])
def test_hp_random_degrees(orc, hp, dim, n, dirichlet):
    # BASELINE config 3 in small: per-element degree 1..6, random
    rng = np.random.default_rng(1887)
    deg = rng.integers(1, 7, int(np.prod(n))).astype(np.int32)
    m = orc.Mesh(n, degree=deg, sigma=2.0, dirichlet=dirichlet)
    x = orc.fill_random(m.ndof)
    ref = m.apply_mf(x, threads=orc.max_threads())
    ctx = hp.Context(n, degree=deg, sigma=2.0, dirichlet=dirichlet)
    assert np.array_equal(ctx.block_offsets(), m.offsets)
    y = hp.Operator(ctx).apply(x)
    assert rel(y, ref) < TOL


def test_hp_single_high_element(orc, hp):
    # matrix-free/test/testsumfactor.cc: first element has degree k+1
    for k in (1, 2, 3, 4):
        deg = np.full(64, k, dtype=np.int32)
        deg[0] = k + 1
        m = orc.Mesh((8, 8), degree=deg)
        x = orc.fill_random(m.ndof)
        y = hp.Operator(hp.Context((8, 8), degree=deg)).apply(x)
        assert rel(y, m.assemble().mv(x)) < TOL


def test_high_order_and_q0(orc, hp):
    for p in (0, 7, 9, 13):
        n = (3, 2, 2) if p < 13 else (2, 2)
        m = orc.Mesh(n, degree=p)
        x = orc.fill_random(m.ndof)
        y = hp.Operator(hp.Context(n, degree=p)).apply(x)
        assert rel(y, m.apply_mf(x, threads=orc.max_threads())) < (1e-11 if p == 13 else TOL)


def test_cfg2_full_size(orc, hp):
    # BASELINE config 2: 64^3 Q3.  Direct comparison with the (threaded) matrix-free oracle plus
    # size-independent properties: symmetry and linearity.
    n = (64, 64, 64)
    m = orc.Mesh(n, degree=3, sigma=2.0, dirichlet=True)
    x = orc.fill_random(m.ndof)
    ctx = hp.Context(n, degree=3, sigma=2.0, dirichlet=True)
    op = hp.Operator(ctx)
    y = op.apply(x)
    ref = m.apply_mf(x, threads=orc.max_threads())
    assert rel(y, ref) < TOL
    z = orc.fill_random(m.ndof, seed=7)
    Az = op.apply(z)
    assert abs(z @ y - x @ Az) <= 1e-11 * abs(z @ y)
    assert rel(op.apply(2.0 * x - 3.0 * z), 2.0 * y - 3.0 * Az) < TOL


@pytest.mark.parametrize("n,p", [((32, 32, 18), 3), ((24, 20, 26), 4), ((40, 40, 37), 2)])
def test_host_pointer_apply_pipelined_slabs(orc, hp, n, p):
    # hpdg_op_apply with host pointers streams z-slabs (H2D / kernel / D2H overlapped); ragged slab heights
    m = orc.Mesh(n, degree=p)
    assert m.ndof >= 1 << 20
    x = orc.fill_random(m.ndof)
    ctx = hp.Context(n, degree=p)
    hx, px = ctx.host_alloc(m.ndof)
    hy, py = ctx.host_alloc(m.ndof)
    hx[:] = x
    hp.Operator(ctx, factor=-2.0).apply(px, py)
    assert rel(hy, -2.0 * m.apply_mf(x, threads=orc.max_threads())) < TOL


def test_device_resident_api(orc, hp):
    n = (6, 6, 6)
    m = orc.Mesh(n, degree=3)
    x = orc.fill_random(m.ndof)
    ctx = hp.Context(n, degree=3)
    dx = ctx.upload(x)
    dy = ctx.vec_alloc()
    before = ctx.launch_count
    hp.Operator(ctx).apply_device(dx, dy)
    assert ctx.launch_count == before + 1
    assert rel(ctx.download(dy), m.apply_mf(x)) < TOL
    assert abs(ctx.dot_device(dx, dx) - x @ x) < 1e-10 * (x @ x)
    ms = ctx.time_apply_device(dx, dy, 3)
    assert ms > 0
    ctx.vec_free(dx)
    ctx.vec_free(dy)


def test_operator_tuple_accumulates(orc, hp):
    # matrix-free/test/testoperator.cc:80-98 restated with the SIPG local operator: a tuple of two local operators with factors
    # 1 and 2 gives (1 + 2) A x; Ax is zeroed once (operator.hh:42) and every local operator adds its part (:44-55).
    # Covers the persistent Q3 kernel, the tile kernel (Q4, ragged Q3) and the generic hp kernel, host and device entry points.
    rng = np.random.default_rng(5)
    for n, deg in [((8, 4, 4), 3), ((5, 3, 2), 3), ((3, 3, 3), 4), ((4, 3, 2), rng.integers(1, 5, 24).astype(np.int32)), ((6, 5), 2)]:
        m = orc.Mesh(n, degree=deg)
        x = orc.fill_random(m.ndof)
        ref = m.apply_mf(x, threads=orc.max_threads())
        ctx = hp.Context(n, degree=deg)
        op = hp.Operator.from_local_operators([hp.IPDGOperator(ctx, factor=1.0), hp.IPDGOperator(ctx, factor=2.0)])
        Ax = np.full(m.ndof, 7.0)  # garbage on entry: apply() overwrites
        op.apply(x, Ax)
        assert rel(Ax, 3.0 * ref) < TOL
        dx, dy = ctx.upload(x), ctx.upload(Ax)
        op.apply_device(dx, dy)
        assert rel(ctx.download(dy), 3.0 * ref) < TOL
        ctx.close()


def test_blas1_device(orc, hp):
    # DynamicBlockVector BLAS-1 (common/dynamicbvector.hh:185-314): =, *=, axpy, dot, two_norm
    n = (5, 4, 3)
    ctx = hp.Context(n, degree=3)
    nd = ctx.dimension()
    x, y = orc.fill_random(nd), orc.fill_random(nd, seed=9)
    dx, dy, dz = ctx.upload(x), ctx.upload(y), ctx.vec_alloc()
    assert abs(ctx.two_norm_device(dx) - np.linalg.norm(x)) < 1e-12 * np.linalg.norm(x)
    ctx.assign_device(dx, dz)
    ctx.scale_device(-0.5, dz)
    ctx.axpy_device(2.0, dy, dz)
    ctx.sync()
    assert rel(ctx.download(dz), -0.5 * x + 2.0 * y) < 1e-15
    assert abs(ctx.dot_device(dx, dy) - x @ y) < 1e-12 * np.linalg.norm(x) * np.linalg.norm(y)
    ctx.close()


@pytest.mark.parametrize("precond", [0, 1, 2])
def test_pcg_solves(orc, hp, precond):
    # the Krylov loop around the hot path: CG / block-Jacobi PCG / V-cycle PCG must all solve A x = b (SPD with Dirichlet faces)
    n = (4, 4, 4)
    m = orc.Mesh(n, degree=4, dirichlet=True)
    xs = orc.fill_random(m.ndof)
    b = m.apply_mf(xs, threads=orc.max_threads())
    ctx = hp.Context(n, degree=4, dirichlet=True)
    ctx.build_p_hierarchy()
    cg = hp.ConjugateGradients(ctx, precond=precond, damping=0.75 if precond == 2 else 1.0, smooth=2, tol=1e-11,
                               maxit=2000, check_every=3 if precond == 0 else 1)
    x = np.zeros(m.ndof)
    its = cg.solve(x, b)
    assert cg.relres <= 1e-11 and its < 2000
    assert rel(m.apply_mf(x, threads=orc.max_threads()), b) < 1e-10
    if precond == 2:
        assert its < 100         # multigrid preconditioning (plain CG needs ~10x more on this mesh)
    ctx.close()


def test_loop_solver_mg(orc, hp):
    # LoopSolver + energy norm around the multigrid step (buildingblocks/solve.hh:150-166), checked against the same loop run on
    # the oracle's V-cycle: identical iteration count and iterate
    n = (4, 4, 4)
    fine = orc.Mesh(n, degree=4)
    l1 = fine.coarsen(2)
    l0 = l1.coarsen(1)
    b = orc.fill_random(fine.ndof)
    x = orc.fill_random(fine.ndof, seed=3) * 0.1
    tol, it_ref = 1e-4, 0
    A = lambda v: fine.apply_mf(v, threads=orc.max_threads())
    xr = x.copy()
    for it_ref in range(1, 31):
        old = xr.copy()
        xr, _ = orc.vcycle([l0, l1, fine], None, xr, b.copy(), smoother=1, damping=0.75)
        c = xr - old
        if np.sqrt(c @ A(c)) / np.sqrt(old @ A(old)) < tol:
            break
    ctx = hp.Context(n, degree=4)
    ctx.build_p_hierarchy()
    mg = hp.Multigrid(ctx, form=hp.JACOBI_FD, damping=0.75)
    solver = hp.LoopSolver(mg, maxIterations=30, tolerance=tol)
    dx, db = ctx.upload(x), ctx.upload(b)
    solver.solve_device(dx, db)
    assert solver.iterationCount() == it_ref
    assert rel(ctx.download(dx), xr) < 1e-9
    assert rel(ctx.download(db), b) == 0.0   # the right-hand side is left alone (mgwrapper.hh:23)
    ctx.close()


def test_two_contexts_two_threads(orc, hp):
    # per-context (= per-device) kernel attributes and scratch: two contexts used from two host threads at once, and -- when the box
    # has a second GPU -- on two devices of one process
    import threading
    import ctypes
    ndev = ctypes.c_int(0)
    try:
        ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(ndev))
    except OSError:
        ndev.value = 1
    n = (8, 8, 4)
    m = orc.Mesh(n, degree=3)
    x = orc.fill_random(m.ndof)
    ref = m.apply_mf(x, threads=orc.max_threads())
    jref = m.blockjacobi_apply(x, factor=0.75)
    errs = []

    def work(dev):
        try:
            ctx = hp.Context(n, degree=3, device=dev)
            for _ in range(5):
                assert rel(hp.Operator(ctx).apply(x), ref) < TOL
                assert rel(hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=0.75)(x), jref) < TOL
                dx = ctx.upload(x)
                assert abs(ctx.dot_device(dx, dx) - x @ x) < 1e-10 * (x @ x)
                ctx.vec_free(dx)
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ths = [threading.Thread(target=work, args=(d % max(ndev.value, 1),)) for d in range(2)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs


@pytest.mark.parametrize("form", [0, 1])
def test_block_jacobi_vs_oracle(orc, hp, form):
    rng = np.random.default_rng(2)
    for n, deg in [((4, 3, 3), rng.integers(1, 5, 36).astype(np.int32)), ((5, 4), rng.integers(1, 7, 20).astype(np.int32)),
                   ((3, 3, 3), 3)]:
        for dirichlet in (True, False):
            m = orc.Mesh(n, degree=deg, dirichlet=dirichlet)
            r = orc.fill_random(m.ndof)
            ref = m.blockjacobi_apply(r, factor=0.75)
            ctx = hp.Context(n, degree=deg, dirichlet=dirichlet)
            jac = hp.BlockJacobi(ctx, form=form, damping=0.75)
            assert rel(jac(r), ref) < 1e-11
            for e in (0, m.nelem // 2, m.nelem - 1):
                assert np.abs(jac.diag_block(e) - m.diag_block_mf(e)).max() < 1e-12 * np.abs(m.diag_block_mf(e)).max()


def test_fd_jacobi_persistent_q3(orc, hp):
    # uniform Q3 bricks whose extents are multiples of 4 run the persistent tile kernel (jacobi_uniform_q3p.cuh); variant 40
    # keeps the one-tile-per-CTA kernel: both must match the oracle's exact block solve (ipdgblockjacobi.hh:58-178)
    for n, L, dirichlet in [((4, 4, 4), None, True), ((8, 4, 12), [1.0, 1.5, 0.5], True), ((8, 8, 8), None, False),
                            ((12, 4, 4), [2.0, 1.0, 1.0], False)]:
        m = orc.Mesh(n, L=L, degree=3, dirichlet=dirichlet)
        r = orc.fill_random(m.ndof)
        ref = m.blockjacobi_apply(r, factor=0.75)
        for variant in (0, 40):
            ctx = hp.Context(n, L=L, degree=3, dirichlet=dirichlet)
            ctx.set_option("variant", variant)
            jac = hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=0.75)
            before = ctx.launch_count
            assert rel(jac(r), ref) < 1e-12
            assert ctx.launch_count == before + 1
            # a second damping on the same level rebuilds the reciprocal table
            assert rel(hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=0.5)(r), ref * (0.5 / 0.75)) < 1e-12
            ctx.close()


def test_fd_jacobi_persistent_q4(orc, hp):
    # jacobi_uniform_q4p.cuh: uniform Q4 bricks with extents multiple of (4, 4, 2) run the persistent kernel; variant 40 keeps the
    # one-tile-per-CTA kernel: both must match the oracle's exact block solve
    for n, L, dirichlet in [((4, 4, 2), None, True), ((8, 4, 6), [1.0, 1.5, 0.5], True), ((8, 8, 4), None, False), ((4, 12, 2), None, False)]:
        m = orc.Mesh(n, L=L, degree=4, dirichlet=dirichlet)
        r = orc.fill_random(m.ndof)
        ref = m.blockjacobi_apply(r, factor=0.75)
        for variant in (0, 40):
            ctx = hp.Context(n, L=L, degree=4, dirichlet=dirichlet)
            ctx.set_option("variant", variant)
            assert rel(hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=0.75)(r), ref) < 1e-12
            ctx.close()


def test_transfer_vs_oracle(orc, hp):
    rng = np.random.default_rng(4)
    for n in [(3, 4, 2), (5, 3)]:
        deg = rng.integers(1, 7, int(np.prod(n))).astype(np.int32)
        deg[0] = 6
        fine = orc.Mesh(n, degree=deg)
        ctx = hp.Context(n, degree=deg)
        nl = ctx.build_p_hierarchy()
        assert nl == 3  # p_max = 6: caps 6/4 = 1, 6/2 = 3 (solversetup.hh:77,94)
        l1 = fine.coarsen(3)
        l0 = l1.coarsen(1)
        assert np.array_equal(ctx.level_degrees(1), l1.degree) and np.array_equal(ctx.level_degrees(0), l0.degree)
        xf = orc.fill_random(fine.ndof)
        t2 = hp.OrderTransfer(ctx, 2)
        xc = t2.restrict(xf)
        assert rel(xc, fine.restrict(l1, xf)) < TOL
        assert rel(t2.prolong(xc), fine.prolong(l1, xc)) < TOL
        t1 = hp.OrderTransfer(ctx, 1)
        xcc = t1.restrict(xc)
        assert rel(xcc, l1.restrict(l0, xc)) < TOL
        assert rel(t1.prolong(xcc), l1.prolong(l0, xcc)) < TOL


@pytest.mark.parametrize("p", [2, 3, 4, 5, 6])
def test_transfer_uniform_levels(orc, hp, p):
    # uniform p-hierarchy (e.g. 4 -> 2 -> 1, the BASELINE config 4 transfer): specialised warp-per-element kernels
    n = (5, 3, 4)
    fine = orc.Mesh(n, degree=p)
    ctx = hp.Context(n, degree=p)
    nl = ctx.build_p_hierarchy()
    levels = [fine]
    for l in range(nl - 2, -1, -1):
        levels.insert(0, levels[0].coarsen(int(ctx.level_degrees(l)[0])))
    for l in range(nl - 1, 0, -1):
        f, c = levels[l], levels[l - 1]
        xf = orc.fill_random(f.ndof, seed=l)
        t = hp.OrderTransfer(ctx, l)
        xc = t.restrict(xf)
        assert rel(xc, f.restrict(c, xf)) < TOL
        assert rel(t.prolong(xc), f.prolong(c, xc)) < TOL


def test_coarse_level_operator_is_galerkin_product(orc, hp):
    # level operators below the finest are T^T A T (ordertransfer.hh:124-144)
    n = (3, 3, 2)
    fine = orc.Mesh(n, degree=4)
    l1 = fine.coarsen(2)
    l0 = l1.coarsen(1)
    Af = fine.assemble()
    A1 = fine.galerkin_restrict(l1, Af)
    A0 = l1.galerkin_restrict(l0, A1)
    ctx = hp.Context(n, degree=4)
    assert ctx.build_p_hierarchy() == 3
    for lvl, mesh, A in ((1, l1, A1), (0, l0, A0)):
        x = orc.fill_random(mesh.ndof)
        assert rel(hp.Operator(ctx, level=lvl).apply(x), A.mv(x)) < TOL
        assert ctx.uses_uniform_kernel(lvl)


@pytest.mark.parametrize("form", [0, 1])
def test_vcycle_vs_oracle(orc, hp, form):
    # BASELINE config 4 in small: Q4 -> Q2 -> Q1, 5 pre + 5 post damped block-Jacobi steps, 5 coarse iterations
    n = (4, 4, 4)
    fine = orc.Mesh(n, degree=4)
    l1 = fine.coarsen(2)
    l0 = l1.coarsen(1)
    b = orc.fill_random(fine.ndof)
    x0 = orc.fill_random(fine.ndof, seed=3) * 0.1
    xr, rr = orc.vcycle([l0, l1, fine], None, x0, b, smoother=1, damping=0.75)
    ctx = hp.Context(n, degree=4)
    ctx.build_p_hierarchy()
    mg = hp.Multigrid(ctx, form=form, damping=0.75)
    x = x0.copy()
    bb = b.copy()
    mg.apply(x, bb)
    assert rel(x, xr) < 1e-11 and rel(bb, rr) < 1e-10
    # the returned b is the residual of the returned x (multigrid_impl.hh:60-61)
    assert rel(bb, b - fine.apply_mf(x, threads=orc.max_threads())) < 1e-10


def test_cfg1_assemble_matvec_blockgs(orc, hp):
    # BASELINE config 1: 2D 16x16 Q2: assemble + matvec + 10 block-GS sweeps (test_dynamicblockgs.cc / testdg.cc)
    m = orc.Mesh((16, 16), degree=2, sigma=2.0, dirichlet=True)
    A = m.assemble()
    rowptr, col, boff, val = A.export()
    ctx = hp.Context((16, 16), degree=2, sigma=2.0, dirichlet=True)
    G = hp.AssembledMatrix(ctx)
    assert np.array_equal(G.block_row_ptr, rowptr) and np.array_equal(G.block_col, col) and np.array_equal(G.block_off, boff)
    assert np.linalg.norm(G.entries - val) < 1e-12 * np.linalg.norm(val)          # test_matrices.cc: Frobenius
    x = m.interpolate_normsq()
    assert rel(G.mv(x), A.mv(x)) < TOL
    b = np.ones(m.ndof)
    xr = np.ones(m.ndof)
    xg = np.ones(m.ndof)
    gs = hp.DynamicBlockGS(G)
    gs.setProblem(xg, b)
    for _ in range(10):
        A.blockgs_iterate(b, xr)
        gs.iterate()
    assert rel(xg, xr) < TOL


@pytest.mark.parametrize("dim,n", [(2, (5, 4)), (3, (3, 4, 2)), (3, (1, 3, 1))])
def test_assembled_hp_and_blockgs(orc, hp, dim, n):
    rng = np.random.default_rng(11)
    deg = rng.integers(1, 5, int(np.prod(n))).astype(np.int32)
    for dirichlet in (True, False):
        m = orc.Mesh(n, L=[1.0, 1.5, 0.75][:dim], degree=deg, dirichlet=dirichlet)
        A = m.assemble()
        rowptr, col, boff, val = A.export()
        ctx = hp.Context(n, L=[1.0, 1.5, 0.75][:dim], degree=deg, dirichlet=dirichlet)
        G = hp.AssembledMatrix(ctx)
        assert np.array_equal(G.block_col, col) and np.array_equal(G.block_off, boff)
        assert np.linalg.norm(G.entries - val) < 1e-12 * np.linalg.norm(val)
        x = orc.fill_random(m.ndof)
        assert rel(G.mv(x), A.mv(x)) < TOL
        b = orc.fill_random(m.ndof, 5)
        xr, xg = x.copy(), x.copy()
        gs = hp.DynamicBlockGS(G)
        gs.setProblem(xg, b)
        for _ in range(3):
            A.blockgs_iterate(b, xr)
            gs.iterate()
        assert rel(xg, xr) < 1e-11


@pytest.mark.parametrize("dim,n", [(2, (6, 5)), (3, (3, 4, 2))])
def test_l1_smoother(orc, hp, dim, n):
    # iterationsteps/l1smoother.hh:20-145 on the device against the oracle: ghost list with a duplicate, hp degrees, 4 sweeps
    rng = np.random.default_rng(7)
    ne = int(np.prod(n))
    deg = rng.integers(1, 4, ne).astype(np.int32)
    m = orc.Mesh(n, degree=deg, sigma=2.0, dirichlet=True)
    A = m.assemble()
    ghosts = [0, n[0] - 1, n[0] - 1, ne - 1, ne // 2]
    reg = A.l1_regularization(ghosts)
    ctx = hp.Context(n, degree=deg, sigma=2.0, dirichlet=True)
    G = hp.AssembledMatrix(ctx)
    b = orc.fill_random(m.ndof)
    x, xr = np.ones(m.ndof), np.ones(m.ndof)
    sm = hp.L1Smoother(ghosts)
    sm.setProblem(G, x, b)
    sm.preprocess()
    for _ in range(4):
        sm.iterate()
        A.l1_iterate(reg, b, xr)
    assert rel(x, xr) < TOL
    # an l1 sweep is not a plain block-GS sweep (the regularisation is active) ...
    xg = np.ones(m.ndof)
    for _ in range(4):
        A.blockgs_iterate(b, xg)
    assert rel(x, xg) > 1e-6
    # ... and iterate() before preprocess() is an error, as is a ghost index outside the matrix
    ctx2 = hp.Context(n, degree=deg)
    G2 = hp.AssembledMatrix(ctx2)
    bad = hp.L1Smoother([0])
    bad.setProblem(G2, np.ones(m.ndof), b)
    with pytest.raises(hp.HpdgError):
        bad.iterate()
    bad2 = hp.L1Smoother([ne])
    bad2.setProblem(G2, np.ones(m.ndof), b)
    with pytest.raises(hp.HpdgError):
        bad2.preprocess()


def test_dynamicblockgs_small_system_on_device(orc, hp):
    # test/test_dynamicblockgs.cc:24-44 on the device: 2x2, k=2, sigma 1.5, b = 1, x0 = 1, 100 sweeps -> |b - Ax| < 1e-13
    ctx = hp.Context((2, 2), degree=2, sigma=1.5, dirichlet=True)
    G = hp.AssembledMatrix(ctx)
    n = ctx.dimension()
    b, x = np.ones(n), np.ones(n)
    gs = hp.DynamicBlockGS(G)
    gs.setProblem(x, b)
    for _ in range(100):
        gs.iterate()
    assert np.linalg.norm(b - G.mv(x)) < 1e-13


def test_vcycle_q3_fine_level_persistent_kernel(orc, hp):
    # Q3 -> Q1 (solversetup.hh:77,94 for pmax = 3) on a mesh whose fine level runs the persistent Q3 kernel: covers its
    # accumulate mode (r -= A c is fused into the operator kernel)
    n = (8, 4, 4)
    fine = orc.Mesh(n, degree=3)
    l0 = fine.coarsen(1)
    b = orc.fill_random(fine.ndof)
    x0 = orc.fill_random(fine.ndof, seed=3) * 0.1
    xr, rr = orc.vcycle([l0, fine], None, x0, b, smoother=1, damping=0.75)
    ctx = hp.Context(n, degree=3)
    ctx.build_p_hierarchy()
    mg = hp.Multigrid(ctx, form=hp.JACOBI_FD, damping=0.75)
    x, bb = x0.copy(), b.copy()
    mg.apply(x, bb)
    assert rel(x, xr) < 1e-11 and rel(bb, rr) < 1e-10


@pytest.mark.parametrize("p,n", [(4, (4, 4, 4)), (3, (8, 4, 4))])
def test_vcycle_xacc_fusion_placements(orc, hp, p, n):
    # x += c of the smoothing steps is fused either into the operator kernel that applies c next (default on uniform levels) or
    # into the block-Jacobi kernel (option xacc_in_apply = 0): both must reproduce the oracle's cycle (multigrid_impl.hh:76-81)
    fine = orc.Mesh(n, degree=p)
    levels = [fine.coarsen(2), fine] if p == 4 else [fine.coarsen(1), fine]
    if p == 4:
        levels = [levels[0].coarsen(1)] + levels
    b = orc.fill_random(fine.ndof)
    x0 = orc.fill_random(fine.ndof, seed=3) * 0.1
    xr, rr = orc.vcycle(levels, None, x0, b, smoother=1, damping=0.75)
    for in_apply in (1, 0):
        ctx = hp.Context(n, degree=p)
        ctx.build_p_hierarchy()
        ctx.set_option("xacc_in_apply", in_apply)
        x, bb = x0.copy(), b.copy()
        hp.Multigrid(ctx, form=hp.JACOBI_FD, damping=0.75).apply(x, bb)
        assert rel(x, xr) < 1e-11 and rel(bb, rr) < 1e-10
        ctx.close()


def test_vcycle_with_reference_default_smoother(orc, hp):
    # the reference's own p-MG configuration (solversetup.hh:139-145,198-215): DynamicBlockGS on the Galerkin level matrices
    n = (3, 3, 3)
    fine = orc.Mesh(n, degree=4)
    l1 = fine.coarsen(2)
    l0 = l1.coarsen(1)
    Af = fine.assemble()
    A1 = fine.galerkin_restrict(l1, Af)
    A0 = l1.galerkin_restrict(l0, A1)
    b = orc.fill_random(fine.ndof)
    x0 = orc.fill_random(fine.ndof, seed=3) * 0.1
    xr, rr = orc.vcycle([l0, l1, fine], [A0, A1, Af], x0, b, smoother=0, damping=1.0)
    ctx = hp.Context(n, degree=4)
    ctx.build_p_hierarchy()
    x, bb = x0.copy(), b.copy()
    hp.Multigrid(ctx, form=hp.SMOOTHER_BLOCKGS, damping=1.0).apply(x, bb)
    assert rel(x, xr) < 1e-11 and rel(bb, rr) < 1e-10
    # and it contracts much faster than damped block Jacobi
    assert np.linalg.norm(bb) < 0.2 * np.linalg.norm(b - Af.mv(x0))


@pytest.mark.parametrize("dim,n,pmax,uniform", [(2, (16, 16), 2, True), (2, (5, 4), 4, False), (3, (3, 4, 2), 4, False),
                                                (3, (8, 4, 4), 3, True), (3, (1, 3, 1), 5, False), (3, (4, 4, 4), 4, True)])
def test_matrix_free_blockgs(orc, hp, dim, n, pmax, uniform):
    # DynamicBlockGS::iterate (iterationsteps/dynamicblockgs.hh:94-126, GSCore :17-40) WITHOUT an assembled matrix, against the
    # oracle's sweep on the assembled matrix: cfg1 (10 sweeps from x = 1, b = 1: test_dynamicblockgs.cc:31-32), hp meshes, 3-D
    rng = np.random.default_rng(13)
    deg = pmax if uniform else rng.integers(1, pmax + 1, int(np.prod(n))).astype(np.int32)
    for dirichlet in (True, False):
        m = orc.Mesh(n, L=[1.0, 1.5, 0.75][:dim], degree=deg, sigma=2.0, dirichlet=dirichlet)
        A = m.assemble()
        ctx = hp.Context(n, L=[1.0, 1.5, 0.75][:dim], degree=deg, sigma=2.0, dirichlet=dirichlet)
        b = np.ones(m.ndof) if uniform else orc.fill_random(m.ndof, 5)
        xr = np.ones(m.ndof) if uniform else orc.fill_random(m.ndof)
        xg = xr.copy()
        gs = hp.MatrixFreeBlockGS(ctx)
        gs.setProblem(xg, b)
        for _ in range(10 if dim == 2 else 3):
            A.blockgs_iterate(b, xr)
            gs.iterate()
        assert rel(xg, xr) < TOL
        ctx.close()


def test_vcycle_with_matrix_free_blockgs(orc, hp):
    # the reference's p-MG configuration (solversetup.hh:139-145,198-215) with the matrix-free DynamicBlockGS sweeps on every
    # level, against the oracle's cycle on the assembled Galerkin matrices
    n = (3, 3, 3)
    fine = orc.Mesh(n, degree=4)
    l1 = fine.coarsen(2)
    l0 = l1.coarsen(1)
    Af = fine.assemble()
    A1 = fine.galerkin_restrict(l1, Af)
    A0 = l1.galerkin_restrict(l0, A1)
    b = orc.fill_random(fine.ndof)
    x0 = orc.fill_random(fine.ndof, seed=3) * 0.1
    xr, rr = orc.vcycle([l0, l1, fine], [A0, A1, Af], x0, b, smoother=0, damping=1.0)
    ctx = hp.Context(n, degree=4)
    ctx.build_p_hierarchy()
    x, bb = x0.copy(), b.copy()
    hp.Multigrid(ctx, form=hp.SMOOTHER_BLOCKGS_MF, damping=1.0).apply(x, bb)
    assert rel(x, xr) < 1e-11 and rel(bb, rr) < 1e-10


def test_matrix_free_blockgs_at_size(hp):
    # a mesh whose assembled matrix would not be built (32^3 Q3: 7.5 GB): the sweeps run and reduce the residual of A x = b
    n = (32, 32, 32)
    ctx = hp.Context(n, degree=3, dirichlet=True)
    nd = ctx.dimension()
    b = np.random.default_rng(3).standard_normal(nd)
    dx, db, dr = ctx.upload(np.zeros(nd)), ctx.upload(b), ctx.vec_alloc()
    gs = hp.MatrixFreeBlockGS(ctx)
    res = []
    for _ in range(3):
        gs.iterate_device(dx, db)
        hp.Operator(ctx).apply_device(dx, dr)
        res.append(np.linalg.norm(b - ctx.download(dr)))
    assert res[0] < np.linalg.norm(b) and res[2] < res[1] < res[0]
    ctx.close()


def test_cfg5_full_size_properties(hp):
    # BASELINE config 5 brick: 128^3 elements, Q4 (262 144 000 DoF, 2.1 GB per vector) -- too large for the CPU oracle, so the
    # size-independent properties of the operator are checked on the device: symmetry, linearity, positivity.
    import torch
    if torch.cuda.mem_get_info()[0] < 16e9:
        pytest.skip("needs 16 GB of free HBM")
    n = (128, 128, 128)
    ctx = hp.Context(n, degree=4, sigma=2.0, dirichlet=True)
    nd = ctx.dimension()
    assert nd == 262144000 and ctx.uses_uniform_kernel()
    rng = np.random.default_rng(5)
    hx, px = ctx.host_alloc(nd)
    hx[:] = rng.standard_normal(nd)
    dx = ctx.upload(px)
    hx[:] = rng.standard_normal(nd)
    dz = ctx.upload(px)
    dAx, dAz, dw = ctx.vec_alloc(), ctx.vec_alloc(), ctx.vec_alloc()
    op = hp.Operator(ctx)
    op.apply_device(dx, dAx)
    op.apply_device(dz, dAz)
    zAx, xAz, xAx = ctx.dot_device(dz, dAx), ctx.dot_device(dx, dAz), ctx.dot_device(dx, dAx)
    assert abs(zAx - xAz) <= 1e-11 * abs(xAx)          # symmetry
    assert xAx > 0                                        # SPD with Dirichlet faces
    # linearity: A(2x - 3z) = 2Ax - 3Az
    ctx.axpy_device(1.0, dx, dx)                          # x <- 2x
    ctx.axpy_device(-3.0, dz, dx)                         # x <- 2x - 3z
    op.apply_device(dx, dw)
    ctx.axpy_device(-2.0, dAx, dw)
    ctx.axpy_device(3.0, dAz, dw)
    ctx.sync()
    assert ctx.dot_device(dw, dw) <= (1e-12 ** 2) * (4 * ctx.dot_device(dAx, dAx) + 9 * ctx.dot_device(dAz, dAz))
    for d in (dx, dz, dAx, dAz, dw):
        ctx.vec_free(d)
    ctx.host_free(px)
    ctx.close()


def test_cfg3_cfg4_full_size_properties(orc, hp):
    # BASELINE configs 3 and 4 at the sizes DESIGN.md quotes numbers for, through size-independent properties.
    # cfg3: 32^3 elements, per-element degree 1..6 (4.3 M DoF): symmetry and positivity of the hp operator and of the block-Jacobi
    # preconditioner (exact element inverses), and parity of the corner element's rows against the oracle (locality).
    rng = np.random.default_rng(1887)
    n = (32, 32, 32)
    deg = rng.integers(1, 7, 32 ** 3).astype(np.int32)
    ctx = hp.Context(n, degree=deg, sigma=2.0, dirichlet=True)
    nd = ctx.dimension()
    x, z = rng.standard_normal(nd), rng.standard_normal(nd)
    op = hp.Operator(ctx)
    Ax, Az = op.apply(x), op.apply(z)
    assert abs(z @ Ax - x @ Az) <= 1e-11 * abs(x @ Ax) and x @ Ax > 0
    jac = hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=1.0)
    Jx, Jz = jac(x), jac(z)
    assert abs(z @ Jx - x @ Jz) <= 1e-11 * abs(x @ Jx) and x @ Jx > 0
    # locality: rows of element (0,0,0) depend only on its face neighbours -- compare with the oracle on the 2x2x2 corner mesh
    # with the same degrees, same spacing, Dirichlet; the corner element's rows agree because its three interior faces see the same
    # neighbours and its three boundary faces are the same Dirichlet faces
    sub = np.array([deg[i + 32 * (j + 32 * k)] for k in range(2) for j in range(2) for i in range(2)], dtype=np.int32)
    m = orc.Mesh((2, 2, 2), L=[2.0 / 32] * 3, degree=sub, sigma=2.0, dirichlet=True)
    off = ctx.block_offsets()
    xs = np.concatenate([x[off[e]:off[e + 1]] for e in (i + 32 * (j + 32 * k) for k in range(2) for j in range(2) for i in range(2))])
    ys = m.apply_mf(xs)
    n0 = (deg[0] + 1) ** 3
    assert rel(Ax[:n0], ys[:n0]) < TOL
    ctx.close()
    # cfg4: 128^3 Q4 -> Q2 -> Q1 V-cycle with damped block Jacobi: the cycle is a contraction in the residual and leaves b = the
    # residual of the returned x (multigrid_impl.hh:60-61)
    import torch
    if torch.cuda.mem_get_info()[0] < 40e9:
        pytest.skip("needs 40 GB of free HBM")
    ctx = hp.Context((128, 128, 128), degree=4, sigma=2.0, dirichlet=True)
    ctx.build_p_hierarchy()
    nd = ctx.dimension()
    hb, pb = ctx.host_alloc(nd)
    hb[:] = np.random.default_rng(3).standard_normal(nd)
    db, dx, dr = ctx.upload(pb), ctx.vec_alloc(), ctx.vec_alloc()
    ctx._ck(hp.lib().hpdg_assign_device(ctx._h, hp.FINEST, db, dx))
    ctx._ck(hp.lib().hpdg_scale_device(ctx._h, hp.FINEST, 0.0, dx))      # x = 0 (vec_alloc does not clear)
    b0 = np.sqrt(ctx.dot_device(db, db))
    mg = hp.Multigrid(ctx, form=hp.JACOBI_FD, damping=0.75)
    mg.apply_device(dx, db)                      # db now holds the residual
    r1 = np.sqrt(ctx.dot_device(db, db))
    hp.Operator(ctx).apply_device(dx, dr)        # A x
    hb2, pb2 = ctx.host_alloc(nd)
    hb2[:] = np.random.default_rng(3).standard_normal(nd)
    db2 = ctx.upload(pb2)
    ctx.axpy_device(-1.0, dr, db2)               # b - A x
    ctx.axpy_device(-1.0, db, db2)               # minus the residual the cycle returned
    ctx.sync()
    assert r1 < 0.8 * b0                          # measured 0.545 for a random right-hand side (p-MG on a fixed mesh, Jacobi smoothing)
    assert np.sqrt(ctx.dot_device(db2, db2)) <= 1e-10 * b0
    ctx.close()


def test_error_behaviour(orc, hp):
    # errors are reported, not aborted on (the reference throws Dune::Exception, e.g. dynamicblockgs.hh:117)
    ctx = hp.Context((4, 4, 4), degree=3)
    x = np.zeros(ctx.dimension())
    with pytest.raises(hp.HpdgError, match="jacobi_setup"):
        ctx._ck(hp.lib().hpdg_jacobi_apply(ctx._h, hp.FINEST, hp.JACOBI_DENSE, x.ctypes.data, x.ctypes.data, 1.0))
    with pytest.raises(hp.HpdgError, match="level index"):
        hp.Operator(ctx, level=3).apply(x)
    with pytest.raises(hp.HpdgError, match="no coarser level"):
        hp.OrderTransfer(ctx, 0).restrict(x)
    with pytest.raises(hp.HpdgError, match="assemble"):
        ctx._ck(hp.lib().hpdg_bcrs_mv(ctx._h, hp.FINEST, x.ctypes.data, x.ctypes.data))
    with pytest.raises(hp.HpdgError, match="unknown option"):
        ctx.set_option("nonsense", 1)

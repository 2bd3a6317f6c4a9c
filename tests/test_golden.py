"""Committed golden fixtures (tests/golden/oracle_golden.npz, made by tests/golden/gen_oracle_golden.py).
CPU: the oracle still reproduces them.  GPU: the CUDA path, through the C ABI, matches them to 1e-12 relative L2."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz"))


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_oracle_reproduces_golden(orc):
    m = orc.Mesh((16, 16), degree=2, sigma=2.0, dirichlet=True)
    assert rel(m.interpolate_normsq(), G["cfg1_x"]) < 1e-15
    assert rel(m.apply_mf(G["cfg1_x"]), G["cfg1_Ax"]) < 1e-14
    m3 = orc.Mesh((5, 6, 3), L=[1.0, 1.5, 0.5], degree=3, sigma=2.0, dirichlet=True)
    assert rel(m3.apply_mf(orc.fill_random(m3.ndof)), G["q3_d_Ax"]) < 1e-14
    mh = orc.Mesh((4, 3, 5), degree=G["hp_deg"], sigma=2.0, dirichlet=True)
    assert rel(mh.apply_mf(orc.fill_random(mh.ndof)), G["hp_Ax"]) < 1e-14
    from oracle import sf2d
    s2 = sf2d.SumFactIPDG2D((7, 6), (1.0, 2.0), G["sf2d_deg"], 2.0, True)
    x2 = orc.fill_random(s2.ndof)
    assert rel(s2.apply(x2), G["sf2d_Ax"]) < 1e-14
    # ... and the quadrature-loop formulation reproduces the sum-factorised one's fixture
    assert rel(orc.Mesh((7, 6), L=[1.0, 2.0], degree=G["sf2d_deg"], sigma=2.0, dirichlet=True).apply_mf(x2), G["sf2d_Ax"]) < 1e-13
    nc = sf2d.RefinedSumFactIPDG2D((5, 4), G["nc_refine"], G["nc_deg"], (1.0, 1.5), 2.0, True)
    assert rel(nc.apply(orc.fill_random(nc.ndof)), G["nc_Ax"]) < 1e-14


@pytest.mark.gpu
def test_cuda_matches_golden(orc, hp):
    ctx = hp.Context((16, 16), degree=2, sigma=2.0, dirichlet=True)
    assert rel(hp.Operator(ctx).apply(G["cfg1_x"]), G["cfg1_Ax"]) < 1e-12
    A = hp.AssembledMatrix(ctx)
    b, x = np.ones(ctx.dimension()), np.ones(ctx.dimension())
    gs = hp.DynamicBlockGS(A)
    gs.setProblem(x, b)
    for _ in range(10):
        gs.iterate()
    assert rel(x, G["cfg1_gs10"]) < 1e-12
    for tag, dirichlet in (("d", True), ("n", False)):
        ctx = hp.Context((5, 6, 3), L=[1.0, 1.5, 0.5], degree=3, sigma=2.0, dirichlet=dirichlet)
        xin = orc.fill_random(ctx.dimension())
        assert rel(hp.Operator(ctx).apply(xin), G[f"q3_{tag}_Ax"]) < 1e-12
        for form in (hp.JACOBI_DENSE, hp.JACOBI_FD):
            assert rel(hp.BlockJacobi(ctx, form=form, damping=0.75)(xin), G[f"q3_{tag}_jac"]) < 1e-11
    ctx = hp.Context((4, 3, 5), degree=G["hp_deg"], sigma=2.0, dirichlet=True)
    xin = orc.fill_random(ctx.dimension())
    assert rel(hp.Operator(ctx).apply(xin), G["hp_Ax"]) < 1e-12
    assert rel(hp.BlockJacobi(ctx, form=hp.JACOBI_FD)(xin), G["hp_jac"]) < 1e-11
    ctx = hp.Context((7, 6), L=[1.0, 2.0], degree=G["sf2d_deg"], sigma=2.0, dirichlet=True)
    assert rel(hp.Operator(ctx).apply(orc.fill_random(ctx.dimension())), G["sf2d_Ax"]) < 1e-12
    ctx = hp.Context.refined_2d((5, 4), G["nc_refine"], G["nc_deg"], L=[1.0, 1.5], sigma=2.0, dirichlet=True)
    assert rel(hp.Operator(ctx).apply(orc.fill_random(ctx.dimension())), G["nc_Ax"]) < 1e-12
    ctx = hp.Context((4, 4, 4), degree=4)
    ctx.build_p_hierarchy()
    xf = orc.fill_random(ctx.dimension())
    t = hp.OrderTransfer(ctx, 2)
    xc = t.restrict(xf)
    assert rel(xc, G["mg_restrict"]) < 1e-12 and rel(t.prolong(xc), G["mg_prolong"]) < 1e-12
    x, b = np.zeros(ctx.dimension()), orc.fill_random(ctx.dimension(), seed=5)
    hp.Multigrid(ctx, form=hp.JACOBI_FD, damping=0.75).apply(x, b)
    assert rel(x, G["mg_vcycle_x"]) < 1e-11 and rel(b, G["mg_vcycle_r"]) < 1e-10

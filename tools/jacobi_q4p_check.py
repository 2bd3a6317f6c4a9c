"""Round-2 starting point: parity + timing of the EXPERIMENTAL persistent Q4 block-Jacobi kernel (jacobi_uniform_q4p.cuh, option
"variant" = 50) against the oracle and the default tile kernel.  Not yet run on a GPU."""
import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/dune-hpdg_b200')
import hpdg_b200 as hp
from oracle import orc
ok = True
for n, L, dirichlet in [((4, 4, 2), None, True), ((8, 4, 6), [1.0, 1.5, 0.5], True), ((8, 8, 4), None, False), ((4, 12, 2), None, False), ((12, 8, 8), None, True)]:
    m = orc.Mesh(n, L=L, degree=4, dirichlet=dirichlet)
    r = orc.fill_random(m.ndof)
    ref = m.blockjacobi_apply(r, factor=0.75)
    for variant in (50, 0):
        ctx = hp.Context(n, L=L, degree=4, dirichlet=dirichlet)
        ctx.set_option("variant", variant)
        c = hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=0.75)(r)
        err = np.linalg.norm(c - ref) / np.linalg.norm(ref)
        ok &= err < 1e-11
        print(n, dirichlet, 'variant', variant, '%.2e' % err, 'OK' if err < 1e-11 else 'FAIL', flush=True)
for variant in (50, 0):
    ctx = hp.Context((64,) * 3, degree=4)
    ctx.set_option("variant", variant)
    nd = ctx.dimension()
    bufs = [(ctx.upload(np.random.default_rng(i).standard_normal(nd)), ctx.vec_alloc()) for i in range(3)]
    jac = hp.BlockJacobi(ctx, form=hp.JACOBI_FD, damping=0.75)
    for i in range(6): jac.apply_device(*bufs[i % 3])
    us = min(sum(jac.time_device(*bufs[i % 3], 20) for i in range(3)) / 3 for _ in range(3)) * 1e3
    print('64^3 Q4 FD block Jacobi, variant %d: %.1f us/apply (CUDA events), %.0f GB/s algorithmic (16 B/DoF)' % (variant, us, nd * 16 / us / 1e3), flush=True)
print('ALL OK' if ok else 'FAILED')

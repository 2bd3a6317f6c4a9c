"""Small driver for ncu: a few operator applies (and optionally Jacobi sweeps) on one workload."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dune-hpdg_b200"))
import numpy as np
import hpdg_b200 as hp

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=64)
ap.add_argument("--p", type=int, default=3)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--jacobi", type=int, default=-1, help="-1 none, 0 dense, 1 fd")
ap.add_argument("--generic", type=int, default=0)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--tune", type=int, default=0)
ap.add_argument("--grid", type=int, default=0, help="persistent Q3 kernel: CTA count (0 = one per SM slot)")
a = ap.parse_args()
ctx = hp.Context((a.n,) * 3, degree=a.p)
ctx.set_option("force_generic", a.generic)
ctx.set_option("variant", a.variant)
ctx.set_option("q3p_grid", a.grid)
ctx.set_option("q3p_tune", a.tune)
nd = ctx.dimension()
x = np.random.default_rng(0).standard_normal(nd)
dx, dy = ctx.upload(x), ctx.vec_alloc()
op = hp.Operator(ctx)
for _ in range(a.reps):
    op.apply_device(dx, dy)
ms = ctx.time_apply_device(dx, dy, 20)
print(f"variant={a.variant} grid={a.grid} tune={a.tune} n={a.n} p={a.p} ndof={nd} apply {ms*1e3:.1f} us  {nd/ms/1e6:.1f} GDoF/s  {16*nd/ms/1e6:.0f} GB/s")
if a.jacobi >= 0:
    jac = hp.BlockJacobi(ctx, form=a.jacobi, damping=0.75)
    for _ in range(a.reps):
        jac.apply_device(dx, dy)
    ms = jac.time_device(dx, dy, 20)
    print(f"variant={a.variant} grid={a.grid} n={a.n} p={a.p} jacobi form {a.jacobi}: {ms*1e3:.1f} us  {nd/ms/1e6:.1f} GDoF/s  {16*nd/ms/1e6:.0f} GB/s (16 B/DoF)")

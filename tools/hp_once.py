import sys
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dune-hpdg_b200'))
import numpy as np, hpdg_b200 as hp
rng = np.random.default_rng(1887)
deg = rng.integers(1, 7, 32 ** 3).astype(np.int32)
ctx = hp.Context((32, 32, 32), degree=deg)
nd = ctx.dimension()
dx, dy = ctx.upload(rng.standard_normal(nd)), ctx.vec_alloc()
op = hp.Operator(ctx)
for _ in range(3):
    op.apply_device(dx, dy)
